import sys, json, time
sys.path.insert(0, "/root/repo")
import bench
from swiftest_b200 import Context
with Context(0) as ctx:
    t0 = time.time()
    r = bench.conservation_reference_test(ctx, int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000)
    print(json.dumps(r, indent=1)); print("wall", time.time() - t0)
