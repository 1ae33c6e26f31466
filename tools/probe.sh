for kw in '{"k":6,"s":2,"c":3,"co":32,"h":128,"w":128,"pad":2}' '{"k":6,"s":2,"c":3,"co":32,"h":128,"w":128,"padding":1,"pad":2}' '{"k":6,"s":2,"c":3,"co":16,"h":130,"w":136,"pad":2}' '{"k":6,"s":2,"c":4,"co":48,"h":128,"w":160,"padding":1,"pad":2}' '{"k":6,"s":2,"c":2,"co":128,"h":160,"w":128,"pad":2}'; do
  echo "== $kw"; timeout 60 python tools/tc_probe.py "$kw" 2>&1 | tail -2 | cut -c1-330
done
