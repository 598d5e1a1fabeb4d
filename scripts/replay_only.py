"""The `replay` key of the bench line alone (frames/s of the offline replay through KinematicICP::RegisterFrame): for quick A/B runs."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

t0 = time.time()
r = bench.run_pipeline_replay(frames=int(sys.argv[1]) if len(sys.argv) > 1 else 24)
r["wall_s"] = time.time() - t0
print(json.dumps(r))
