#!/usr/bin/env python3
"""Pack the reference's motion-clip CSV tables into one .npz (data, not code; float64 exactly as pandas parses them).

Input : /root/reference/Joints CSV With Hand/motion{02_04,08_03,09_03,13_13}{JointPosRad,JointPosRadRelative,
        JointSpeedRadSec,JointVecFromHip}.csv      (read by low_level_env.py:58-71, hier_env.py:61-80)
Output: imitation-learning-rl_b200/data/clips.npz  keys <clip>_{pos,rel,vel,ep}, column order = CSV header order.
"""
import os, sys
import numpy as np
import pandas as pd

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/Joints CSV With Hand"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "imitation-learning-rl_b200", "data", "clips.npz")
CLIPS = ["motion02_04", "motion08_03", "motion09_03", "motion13_13"]
TABLES = dict(pos="JointPosRad", rel="JointPosRadRelative", vel="JointSpeedRadSec", ep="JointVecFromHip")
JCOLS = ["rightHipX", "rightHipY", "rightHipZ", "rightKnee", "leftHipX", "leftHipY", "leftHipZ", "leftKnee",
         "rightShoulderX", "rightShoulderY", "rightElbow", "leftShoulderX", "leftShoulderY", "leftElbow"]
EPCOLS = ["%s_%sposition" % (j, a) for j in ["LeftLeg", "LeftFoot", "RightLeg", "RightFoot", "Head", "LeftForeArm",
                                             "LeftHand", "RightForeArm", "RightHand"] for a in "XYZ"]
out = {}
for c in CLIPS:
    for k, t in TABLES.items():
        df = pd.read_csv(os.path.join(SRC, "%s%s.csv" % (c, t)))
        assert list(df.columns) == (EPCOLS if k == "ep" else JCOLS), (c, t)
        out["%s_%s" % (c, k)] = df.to_numpy(dtype=np.float64)
        print(c, k, out["%s_%s" % (c, k)].shape)
os.makedirs(os.path.dirname(OUT), exist_ok=True)
np.savez_compressed(OUT, **out)
print("wrote", OUT, os.path.getsize(OUT), "bytes")
