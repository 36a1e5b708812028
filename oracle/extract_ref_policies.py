#!/usr/bin/env python3
"""Extract, from the reference repo's own training artefacts (build container only), what the policy-transfer check
needs (tools/policy_transfer.py, tests/test_gpu_transfer.py) and commit it as a fixture:
  * the weights of the low-level policies the reference trained IN PYBULLET and shipped under Log/ (RLlib 1.2 TF
    checkpoints: plain numpy arrays inside two nested pickles; ray itself is not needed to read them),
  * the outcome of the reference's own evaluation of exactly those checkpoints in PyBullet (REF env_check.py writes
    Log/data_<experiment timestamp>_<checkpoint>.json: survival time and mean drift, 36 turning angles x 10 episodes).
These are DATA files of the reference (weights, logs), not source."""
import io
import json
import os
import pickle
import sys

import numpy as np

REF = os.environ.get("ILRL_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("ILRL_GOLDEN_OUT") or os.path.join(ROOT, "tests", "golden")


class _Stub:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, s):
        self.__dict__["state"] = s


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):   # ray / tensorflow classes are absent here: stand-ins with the same name
        try:
            return super().find_class(module, name)
        except Exception:
            return type(name, (_Stub,), {})


RUNS = {   # experiment id -> (checkpoint, evaluation log, what Log/catatan_low_level.txt:596-640 records about the run)
    "PPO_HumanoidBulletEnv-v0-Low_6d114_00000_0_2021-04-30_23-26-25": (1690, "data_2021-04-30_23-26-25_1690.json"),
    "PPO_HumanoidBulletEnv-v0-Low_68eec_00000_0_2021-05-01_23-32-16": (1610, "data_2021-05-01_23-32-16_1610.json"),
}


def main():
    out = {}
    for exp, (ck, log) in RUNS.items():
        path = os.path.join(REF, "Log", "Best Model", "Low", exp, "checkpoint_%d" % ck, "checkpoint-%d" % ck)
        d = _Unpickler(open(path, "rb")).load()
        w = _Unpickler(io.BytesIO(d["worker"])).load()["state"]["default_policy"]
        tag = exp.split("_")[2]
        for k in ("log_std", "fc_1/kernel", "fc_1/bias", "fc_2/kernel", "fc_2/bias", "fc_out/kernel", "fc_out/bias"):
            out["%s/%s" % (tag, k)] = np.asarray(w["default_policy/" + k], dtype=np.float32)
        ev = json.load(open(os.path.join(REF, "Log", log)))
        degs = sorted(int(k) for k in ev)
        out["%s/eval_deg" % tag] = np.array(degs, dtype=np.int32)
        out["%s/eval_timestep" % tag] = np.array([ev[str(g)]["timestep"] for g in degs], dtype=np.int32)
        out["%s/eval_drift" % tag] = np.array([ev[str(g)]["drift"] for g in degs], dtype=np.float64)
        print(tag, "obs", out["%s/fc_1/kernel" % tag].shape[0], "eval", out["%s/eval_timestep" % tag].shape,
              "mean survival %.0f" % out["%s/eval_timestep" % tag].mean())
    np.savez_compressed(os.path.join(OUT, "ref_policies.npz"), **out)


if __name__ == "__main__":
    sys.exit(main())
