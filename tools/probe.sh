for kw in '{"k":3,"s":1,"c":32,"co":64,"h":20,"w":20}' '{"k":3,"s":1,"c":64,"co":32,"h":20,"w":20}' '{"k":3,"s":2,"c":64,"co":128,"h":40,"w":40}'; do
  echo "== $kw"; timeout 60 python tools/tc_probe.py "$kw" 2>&1 | tail -2
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "micro or generated" 2>&1 | tail -15
