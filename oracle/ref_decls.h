/* oracle/ref_decls.h -- TEST INFRASTRUCTURE: prototypes force-included into the
 * reference TUs by build_ref.sh (-include), so the quiet-printf redirection and
 * the arena-size hook are declared before use. */
#ifndef ORACLE_REF_DECLS_H
#define ORACLE_REF_DECLS_H
#include <stddef.h>
#include <stdio.h>
#ifdef __cplusplus
extern "C" {
#endif
int oracle_ref_printf(const char *fmt, ...);
int oracle_ref_fprintf(FILE *f, const char *fmt, ...);
size_t oracle_ref_ddr_size(void);
#ifdef __cplusplus
}
#endif
#endif
