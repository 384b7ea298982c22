python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for pdl in 0 1; do for cfg in "8 4" "8 2"; do set -- $cfg; echo "PDL $pdl"; SEPFWI_PDL=$pdl SEPFWI_LZ=$1 SEPFWI_LZE=$2 python tools/t.py c2 400 1; done; done
for pdl in 0 1; do echo "PDL $pdl"; SEPFWI_PDL=$pdl python tools/quick_perf.py 0 401; done
