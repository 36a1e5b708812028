#!/usr/bin/env python3
"""Re-validate the restated dynamics against a REAL PyBullet, the day one is importable.

Why this exists: the reference's physics lives in the un-vendored `pybullet` / `pybullet_envs` packages, which are not
installable in the build image ("DYNAMICS: PARITY UNPINNED").  Everything Bullet-side in this
repo (csrc/ilrl_constants.h, tools/gen_model.py, oracle/ilrl_oracle.c) was restated from memory of the upstream
source.  This script REPORTS (it does not assert) how far one env step of the oracle — and of the CUDA path when a
GPU is present — is from PyBullet's from identical states, and compares episode-length / return distributions, so
that every recalled constant can be settled in one place.

It has NOT been executed in the build image (no pybullet there).  Usage, on a machine with pybullet:

    python tools/validate_against_pybullet.py --xml /path/to/humanoid_symmetric_2.xml [--states 200] [--gpu]

What it sets up mirrors pybullet_envs (scene_abstract.World.clean_everything / MJCFBasedRobot.reset /
StadiumScene.episode_restart), as recalled in SURVEY.md section 2.3:
    gravity -9.8, setDefaultContactERP(0.9), fixedTimeStep 0.0165, numSolverIterations 5, numSubSteps 4,
    deterministicOverlappingPairs 1, plane with lateralFriction 0.8 / restitution 0.5,
    loadMJCF(flags = URDF_USE_SELF_COLLISION | URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS),
    every joint motor disabled (POSITION_CONTROL, force 0), actions applied as TORQUE_CONTROL.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--xml", required=True, help="humanoid_symmetric_2.xml of the reference repository")
    ap.add_argument("--states", type=int, default=200)
    ap.add_argument("--gpu", action="store_true", help="also compare the CUDA path (needs a B200 and the built library)")
    args = ap.parse_args()
    try:
        import pybullet as p
        import pybullet_data
    except ImportError:
        print("pybullet is not importable here: nothing to validate against (this is the situation DESIGN.md "
              "section 4 describes).")
        return 0
    from oracle import oracle as O
    from oracle.gen_golden import random_phys

    cid = p.connect(p.DIRECT)
    p.resetSimulation(physicsClientId=cid)
    p.setGravity(0, 0, -9.8)
    p.setDefaultContactERP(0.9)
    p.setPhysicsEngineParameter(fixedTimeStep=0.0165, numSolverIterations=5, numSubSteps=4,
                                deterministicOverlappingPairs=1)
    plane = p.loadSDF(os.path.join(pybullet_data.getDataPath(), "plane_stadium.sdf"))
    for b in plane:
        p.changeDynamics(b, -1, lateralFriction=0.8, restitution=0.5)
    flags = p.URDF_USE_SELF_COLLISION | p.URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS
    robot = p.loadMJCF(args.xml, flags=flags)[0]
    model = O.load_model()
    # map our joint order (= pybullet_envs ordered_joints) to Bullet joint indices by name
    name_to_index = {}
    for j in range(p.getNumJoints(robot)):
        info = p.getJointInfo(robot, j)
        name_to_index[info[1].decode()] = j
        p.setJointMotorControl2(robot, j, p.POSITION_CONTROL, positionGain=0.1, velocityGain=0.1, force=0)
    jidx = [name_to_index[n] for n in model["joint_name"]]
    print("Bullet joint info (damping, friction) of the 17 hinges — settles whether MJCF damping/armature are honoured:")
    for n, j in zip(model["joint_name"], jidx):
        info = p.getJointInfo(robot, j)
        print("  %-18s damping %.4g friction %.4g limits [%.4f, %.4f]" % (n, info[6], info[7], info[8], info[9]))
    print("Bullet link masses / local inertia diagonals vs the restated model:")
    for b in range(-1, p.getNumJoints(robot)):
        d = p.getDynamicsInfo(robot, b)
        if d[0] > 0:
            print("  link %3d mass %.6f inertia %s" % (b, d[0], np.round(d[2], 6)))
    print("  restated body masses:", np.round(model["body_mass"], 6))

    def set_state(ph):
        p.resetBasePositionAndOrientation(robot, ph[0:3], ph[3:7])
        p.resetBaseVelocity(robot, ph[7:10], ph[10:13])
        for k, j in enumerate(jidx):
            p.resetJointState(robot, j, ph[13 + k], ph[30 + k])

    def get_state():
        pos, quat = p.getBasePositionAndOrientation(robot)
        lin, ang = p.getBaseVelocity(robot)
        js = [p.getJointState(robot, j) for j in jidx]
        return np.concatenate([pos, quat, lin, ang, [s[0] for s in js], [s[1] for s in js]])

    rng = np.random.default_rng(0)
    worst = dict(q=0.0, qd=0.0, pos=0.0, quat=0.0)
    gpu_env = None
    if args.gpu:
        import ilrl_b200
        gpu_env = ilrl_b200.BatchedHumanoidEnv(1, "low", auto_reset=False)
        worst_gpu = dict(q=0.0, qd=0.0, pos=0.0)
    for s in range(args.states):
        ph = random_phys(rng)
        tau = rng.uniform(-40, 40, 17) * (rng.uniform() < 0.7)
        set_state(ph)
        for k, j in enumerate(jidx):
            p.setJointMotorControl2(robot, j, p.TORQUE_CONTROL, force=float(tau[k]))
        p.stepSimulation()
        got = get_state()
        want = O.physics_step(ph, tau)
        worst["q"] = max(worst["q"], np.abs(got[13:30] - want[13:30]).max())
        worst["qd"] = max(worst["qd"], np.abs(got[30:47] - want[30:47]).max())
        worst["pos"] = max(worst["pos"], np.abs(got[0:3] - want[0:3]).max())
        worst["quat"] = max(worst["quat"], min(np.abs(got[3:7] - want[3:7]).max(), np.abs(got[3:7] + want[3:7]).max()))
        if gpu_env is not None:
            gpu_env.set_state(ph.astype(np.float32)[None, :], None)
            gpu_env.physics_only(tau.astype(np.float32)[None, :])
            g = gpu_env.get_state()[0][0].cpu().numpy()
            worst_gpu["q"] = max(worst_gpu["q"], np.abs(got[13:30] - g[13:30]).max())
            worst_gpu["qd"] = max(worst_gpu["qd"], np.abs(got[30:47] - g[30:47]).max())
            worst_gpu["pos"] = max(worst_gpu["pos"], np.abs(got[0:3] - g[0:3]).max())
    print("one env step (4 substeps) from %d identical random states, max |PyBullet - oracle|:" % args.states, worst)
    if gpu_env is not None:
        print("  and max |PyBullet - CUDA path|:", worst_gpu)
    print("(tolerances the CUDA path meets against the oracle: contact-free q 1e-4 rad, qd 1e-2 rad/s, pos 1e-4 m; "
          "10x looser with active rows — tests/test_gpu_parity.py)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
