"""Per-call latency of PhyloKernel.kernel(tree object, tree object) -- the call kamphir.py:290-291 makes -- at 100, 200
and 500 tips (the same measurement bench.py reports under `kernel_call`)."""
import json
import sys

sys.path.insert(0, ".")
import bench  # noqa: E402

if __name__ == "__main__":
    print(json.dumps(bench.kernel_call_latency(int(sys.argv[1]) if len(sys.argv) > 1 else 64), indent=1))
