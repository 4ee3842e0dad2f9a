for t in 128 64; do echo "== RERANK_THREADS $t"; KEDS_RERANK_THREADS=$t timeout 300 python scripts/perf_configs.py cfg1 2>&1 | python scripts/_pp_perf.py | grep -v pdl0 | head -3; done
KEDS_RERANK_THREADS=64 timeout 600 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
