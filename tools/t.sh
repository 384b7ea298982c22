python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for force in 0 1 2; do SEPFWI_FORCE=$force python tools/t.py c2 400 1; done
for cfg in "8 4" "8 2" "14 4" "14 8" "20 8"; do set -- $cfg; SEPFWI_LZ=$1 SEPFWI_LZE=$2 python tools/t.py c2 400 1; done
python tools/t.py c2 400 8
for force in 0 1 2; do SEPFWI_FORCE=$force python tools/t.py c5s 60 1; done
for lz in 38 62 92; do SEPFWI_LZ=$lz python tools/t.py c5s 60 1; done
