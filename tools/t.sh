python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/quick_perf.py 0 401
for force in 1 2; do SEPFWI_FORCE=$force python tools/t2.py c5s 40 1; done
for force in 1 2; do SEPFWI_FORCE=$force python tools/t2.py c3 200 1; done
