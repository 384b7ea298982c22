for force in 0 1 2; do SEPFWI_FORCE=$force python tools/t2.py c5s 40 1; done
for force in 0 1 2; do SEPFWI_FORCE=$force python tools/t2.py c3 200 1; done
