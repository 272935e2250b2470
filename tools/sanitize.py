import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import torch
from quaternion_mpc_b200 import QuatMpc, abi
from quaternion_mpc_b200.workloads import random_batch, random_gait_states, predict_schedule_numpy
B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
mpc = QuatMpc(horizon=10, max_batch=B)
p = random_batch(B, seed=3, gait="trot")
d = mpc.to_device(p)
r = mpc.grf_update_device(d); torch.cuda.synchronize()
sched = predict_schedule_numpy(random_gait_states(B, seed=3), 10, mpc.cfg.dt)
w = mpc.alloc_warm(B)
for _ in range(2):
    r2 = mpc.grf_update_warm_device(d, w, d_sched=mpc.schedule_to_device(sched)); torch.cuda.synchronize()
g = torch.from_numpy(random_gait_states(B, seed=1).view(np.uint8).reshape(B, -1)).cuda()
s = mpc.predict_contact_schedule(g)
q = torch.rand((B, 12), dtype=torch.float64, device="cuda")
foot, jac = mpc.leg_kinematics(q)
tau = mpc.joint_torques(r, jac, None)
st = mpc.alloc_goal_state()
gin = np.zeros(B, dtype=abi.GOAL_INPUT_DTYPE); gin["torso_quat"][:, 0] = 1
mpc.goal_update(st, torch.from_numpy(gin.view(np.uint8).reshape(B, -1)).cuda(), d)
torch.cuda.synchronize()
print("ok", mpc.results_to_numpy(r)["status"][:8], mpc.results_to_numpy(r2)["iterations"][:8])
# round 2: ConvexMpc on the cooperative kernel, the phased launches, the gait-FSM kernels
from quaternion_mpc_b200 import ConvexMpc
from quaternion_mpc_b200.workloads import random_convex_batch
c = ConvexMpc(horizon=10, max_batch=B)
rc = c.grf_update_device(c.to_device(random_convex_batch(B, seed=3))); torch.cuda.synchronize()
ph = QuatMpc(horizon=10, max_batch=B, kernel="phased")
d2 = mpc.to_device(p)                             # `d` was overwritten in place by goal_update above
rp = ph.grf_update_device(d2); torch.cuda.synchronize()
assert ph.results_to_numpy(rp).tobytes() == mpc.results_to_numpy(r).tobytes()
m20 = QuatMpc(horizon=20, max_batch=B)            # linearisation blocks staged per knot (not shared-memory residents)
r20 = m20.grf_update_device(d2); torch.cuda.synchronize()
fsm = mpc.alloc_leg_fsm()
fin = np.zeros(B, dtype=abi.FOOT_UPDATE_INPUT_DTYPE); fin["movement_mode"] = 1
for _ in range(3):
    fo = mpc.foot_update(fsm, torch.from_numpy(fin.view(np.uint8).reshape(B, -1)).cuda(), 2.2, 0.005, d, g)
torch.cuda.synchronize()
print("ok round 2", c.results_to_numpy(rc)["iterations"][:4], m20.results_to_numpy(r20)["iterations"][:4])
