"""Regenerates tests/golden/*.json from the reference's own golden trajectories.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The fixtures are DATA copied from the reference's test artefacts (not source code):
  legged_ctrl/src/test/test_altro/quat_mpc_test.json        (output of TestAltroQuatMpc.cpp)
  legged_ctrl/src/test/test_altro/trot_quat_mpc_test.json   (output of TestAltroTrotQuatMpc.cpp)
They were produced by the real ALTRO fork (zixinz990/altro@b47202ff) and are the only vectors in
the tree that pin the quaternion-MPC solve.  convex_mpc.json is stale (30 knots vs N=10, SURVEY.md
section 4) and is deliberately not used.
"""
import json
import os

REF = "/root/reference/legged_ctrl/src/test/test_altro"
HERE = os.path.dirname(os.path.abspath(__file__))

for name in ("quat_mpc_test.json", "trot_quat_mpc_test.json"):
    with open(os.path.join(REF, name)) as f:
        d = json.load(f)
    out = {
        "source": f"legged_ctrl/src/test/test_altro/{name}",
        "reference_commit": "4ccae038",
        "state_trajectory": d["state_trajectory"],
        "input_trajectory": d["input_trajectory"],
        "reference_state": d["reference_state"],
        "reference_input": d["reference_input"],
    }
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", name)
