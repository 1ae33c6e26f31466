"""oracle/nna_layout.py -- CPU restatement of the reference's NNA-native layout packers (TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this; the product path never does).

Follows, index for index:
  pack_weights_nmhwsoib2      /root/reference/mars-compiler/src/mars_format.rs:436-470   (size: :472-476)
  convert_nchw_to_ndhwc32     /root/reference/mars-compiler/src/mars_format.rs:490-531   (size: :485-488)
The reference's packers are Rust (no Rust toolchain in this image), so they cannot be run here; the restatement is pinned
instead to the reference's own PYTHON unpackers, which can be imported in the build container
(mgk-decompiler/mgk_decompiler.py:470-540 and mgk-decompiler/scripts/extract_weights_nmhwsoib2.py:52-80): for every fixture
shape, reference_unpack(oracle_pack(w)) == w and the padding bytes are zero -- tests/golden/make_nna_layout_golden.py writes
tests/golden/nna_layout.npz, tests/test_nna_layout.py checks it.  The NDHWC32 converter has no counterpart on the reference's
Python side: it is pinned by its loop restatement (pack_ndhwc32_loops, a literal transcription of :505-527) only.
"""
import numpy as np


def nmhwsoib2_size(out_ch, in_ch, kh, kw):
    return ((out_ch + 31) // 32) * ((in_ch + 31) // 32) * kh * kw * 1024


def ndhwc32_size(batch, channels, height, width):
    return batch * ((channels + 31) // 32) * height * width * 32


def pack_nmhwsoib2_loops(w):
    """literal transcription of mars_format.rs:452-467 (small shapes only)"""
    out_ch, in_ch, kh, kw = w.shape
    m_ifp = (in_ch + 31) // 32
    flat = w.reshape(-1).view(np.uint8)
    packed = np.zeros(nmhwsoib2_size(out_ch, in_ch, kh, kw), np.uint8)
    for o in range(out_ch):
        for i in range(in_ch):
            for h in range(kh):
                for x in range(kw):
                    src = ((o * in_ch + i) * kh + h) * kw + x
                    dst = ((((o // 32 * m_ifp + i // 32) * kh + h) * kw + x) * 32 + o % 32) * 32 + i % 32
                    packed[dst] = flat[src]
    return packed


def pack_nmhwsoib2(w):
    """the same map, vectorised: OIHW int8 [Co, Ci, KH, KW] -> bytes [N_OFP, M_IFP, KH, KW, 32, 32]"""
    w = np.ascontiguousarray(w, dtype=np.int8)
    out_ch, in_ch, kh, kw = w.shape
    n, m = (out_ch + 31) // 32, (in_ch + 31) // 32
    pad = np.zeros((n * 32, m * 32, kh, kw), np.int8)
    pad[:out_ch, :in_ch] = w
    return np.ascontiguousarray(pad.reshape(n, 32, m, 32, kh, kw).transpose(0, 2, 4, 5, 1, 3)).reshape(-1).view(np.uint8)


def unpack_nmhwsoib2(packed, out_ch, in_ch, kh, kw):
    n, m = (out_ch + 31) // 32, (in_ch + 31) // 32
    blocks = np.asarray(packed, dtype=np.uint8)[: n * m * kh * kw * 1024].view(np.int8).reshape(n, m, kh, kw, 32, 32)
    return np.ascontiguousarray(blocks.transpose(0, 4, 1, 5, 2, 3).reshape(n * 32, m * 32, kh, kw)[:out_ch, :in_ch])


def pack_ndhwc32_loops(x):
    """literal transcription of mars_format.rs:505-527 (small shapes only)"""
    batch, ch, hh, ww = x.shape
    d_c32 = (ch + 31) // 32
    flat = x.reshape(-1)
    out = np.zeros(ndhwc32_size(batch, ch, hh, ww), np.uint8)
    for n in range(batch):
        for c in range(ch):
            for h in range(hh):
                for w in range(ww):
                    src = ((n * ch + c) * hh + h) * ww + w
                    dst = (((n * d_c32 + c // 32) * hh + h) * ww + w) * 32 + c % 32
                    out[dst] = flat[src]
    return out


def pack_ndhwc32(x):
    """NCHW uint8 [N, C, H, W] -> bytes [N, D_C32, H, W, 32]"""
    x = np.ascontiguousarray(x, dtype=np.uint8)
    batch, ch, hh, ww = x.shape
    d = (ch + 31) // 32
    pad = np.zeros((batch, d * 32, hh, ww), np.uint8)
    pad[:, :ch] = x
    return np.ascontiguousarray(pad.reshape(batch, d, 32, hh, ww).transpose(0, 1, 3, 4, 2)).reshape(-1)


def unpack_ndhwc32(native, batch, ch, hh, ww):
    d = (ch + 31) // 32
    v = np.asarray(native, dtype=np.uint8)[: batch * d * hh * ww * 32].reshape(batch, d, hh, ww, 32)
    return np.ascontiguousarray(v.transpose(0, 1, 4, 2, 3).reshape(batch, d * 32, hh, ww)[:, :ch])
