timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^E |^FAILED" | head
python scratch/host_profile.py 2>&1 | grep -v "^$" | head -24
