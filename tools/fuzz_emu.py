#!/usr/bin/env python
"""Fuzz of the EMULATED Panda kernel (tools/emu/libb2env_emu.so, build it with tools/emu/build.sh) against the oracle on
contact-rich states (robot-cube / robot-table contacts, coupled islands, up to 47 rows, sweep-capped systems): every step
restarts from the oracle state; contact keys / counts must be exact, converged environments within the single-step tolerances.
    python tools/fuzz_emu.py SEED BATCH STEPS [free_running=0]
Round 1: seeds 1..6 x 192 envs x 20 steps (23 k contact-rich env-steps): no key / row-count mismatch, converged environments
within |dq| 2.3e-6, |dqd| 5.5e-4, cube pose 2.0e-5.
Round 2 (N1 contact model, final kernels): seeds 20..27 x 192 envs x 20 steps (30.7 k contact-rich env-steps, 11655 of them
sweep-capped): no mismatch, converged environments within |dq| 3.4e-06, |dqd| 8.2e-04, cube pose 5.1e-05."""
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import sys,time
sys.path.insert(0,os.path.join(ROOT,'tests')); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'pybullet-robot-envs_b200'))
import numpy as np
from common import *
import test_gpu_parity as T
from oracle import b2oracle
from pybullet_robot_envs.b2env import binding
from pybullet_robot_envs.b2env.binding import B2Sim
lib=binding.load_library(os.path.join(ROOT,'tools','emu','libb2env_emu.so'))
seed=int(sys.argv[1]); B=int(sys.argv[2]); steps=int(sys.argv[3]); free=int(sys.argv[4]) if len(sys.argv)>4 else 0
m,p=panda_task_setup(TASK_PUSH)
orc=b2oracle.Oracle(m,p,B,nthreads=4)
sim=B2Sim(m,p,B,0,lib=lib)
qs,poses=T._contact_rich_states(b2oracle,m,p,B,seed)
tg=targets_for(poses)+np.array([0.3,0,0],np.float32)
orc.reset(poses,tg); orc.state["q"][:]=qs; orc.state["mtarget"][:]=qs
rng=np.random.RandomState(seed+100)
worst=dict(dq=0,dv=0,dc=0); t0=time.time(); capped=0; rowsmax=0; bad=0
copy_state_to_gpu(orc,sim)
for i in range(steps):
    if not free: copy_state_to_gpu(orc,sim)
    a=rng.uniform(-1,1,(B,7)).astype(np.float32)
    o=orc.step(a,1,0); g=sim.step_host(a,1,0)
    gs,os_=sim.get("status"),orc.state["status"]
    if not free:
        if not (gs[:,2:]==os_[:,2:]).all(): bad+=1; print("row/contact count mismatch at step",i,np.argwhere(gs[:,2:]!=os_[:,2:])[:5])
        if not (sim.get("cache_key")==orc.state["cache_key"]).all(): bad+=1; print("key mismatch step",i)
    conv=os_[:,1]<150
    capped+=int((~conv).sum()); rowsmax=max(rowsmax,int(os_[:,3].max()))
    dq=np.abs(sim.get("q")-orc.state["q"]).max(axis=1); dv=np.abs(sim.get("qd")-orc.state["qd"]).max(axis=1); dc=np.abs(sim.get("obj_pose")-orc.state["obj_pose"]).max(axis=1)
    if conv.any():
        worst['dq']=max(worst['dq'],dq[conv].max()); worst['dv']=max(worst['dv'],dv[conv].max()); worst['dc']=max(worst['dc'],dc[conv].max())
    assert (gs[:,0]&1).sum()==0 and np.isfinite(sim.get("obj_pose")).all()
print("seed",seed,"B",B,"steps",steps,"free" if free else "restart","worst(converged)",worst,"capped env-steps",capped,"rows max",rowsmax,"mismatches",bad,"%.0fs"%(time.time()-t0))
