# native ingest on the GPU box's host: the inflate buffer parked between calls against a fresh mapping per call, then the
# thread sweep with the library's own inflate and with zlib's
for park in 1 0 1 0; do LVC_INGEST_PARK=$park python tools/bench_ingest.py --pairs 996767 --threads 16 --reps 6 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('park $park', d['seconds'])"; done
