for mode in 1 2; do for kw in '{"k":3,"s":1,"c":32,"co":32,"h":20,"w":20}' '{"k":3,"s":1,"c":64,"co":64,"h":20,"w":20}' '{"k":3,"s":1,"c":128,"co":32,"h":20,"w":20}' '{"k":3,"s":1,"c":128,"co":32,"h":6,"w":14}'; do
  echo "== halo=$mode $kw"; MARS_TC_HALO=$mode timeout 60 python tools/tc_probe.py "$kw" 2>&1 | tail -1
done; done
