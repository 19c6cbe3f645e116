timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^E |^FAILED" | head
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-200
