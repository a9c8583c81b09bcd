for q in 128 37888 18944 75776; do echo "== quant $q"; FQ_CHUNK_QUANT=$q python tools/pageable_check.py > /tmp/q.txt 2>&1; head -n 2 /tmp/q.txt; done
for s in 75776; do echo "== FQ_DH_CHUNK_ROWS=$((s*8)) quant 37888"; FQ_DH_CHUNK_ROWS=$((s*8)) FQ_CHUNK_QUANT=37888 python tools/pageable_check.py > /tmp/q.txt 2>&1; head -n 2 /tmp/q.txt; done
