for l in d e; do echo "lib $l"; SEPFWI_LIB=$PWD/tmp_libs/$l.so python tools/t2.py c5s 40 1; SEPFWI_LIB=$PWD/tmp_libs/$l.so python tools/t2.py c3 200 1;  done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/quick_perf.py 0 401
