"""Parity bug hunt: the randomized case of tests/test_gpu_fuzz.py (engine vs oracle: forests, scores, votes in all modes
incl. NaN cells and dropped columns) over a range of seeds on the real engine.
    python profiles/fuzz_sweep.py [first] [count] [wide]"""
import sys
import time
import traceback

sys.path.insert(0, ".")
from tests.test_gpu_fuzz import run_case

first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 300
wide = len(sys.argv) > 3 and sys.argv[3] == "wide"
bad = []
t0 = time.time()
for seed in range(first, first + count):
    try:
        run_case(seed, wide)
    except BaseException as e:  # pytest.raises failures included
        bad.append(seed)
        print("SEED", seed, "FAILED:", type(e).__name__, str(e)[:300])
        traceback.print_exc(limit=3)
print(f"{count} {'wide ' if wide else ''}seeds from {first}: {len(bad)} failures {bad} in {time.time() - t0:.0f} s")
