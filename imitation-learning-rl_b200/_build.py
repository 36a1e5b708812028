"""Builds libilrl_b200.so in-tree with nvcc for sm_100a (the only target; there is no CPU build of the product)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libilrl_b200.so")
SRC = os.path.join(HERE, "csrc", "ilrl_capi.cu")
SRC_POLICY = os.path.join(HERE, "csrc", "ilrl_policy.cu")   # tcgen05 fused policy / value forward (rollout collection)
DEPS = [SRC, SRC_POLICY] + [os.path.join(HERE, "csrc", f) for f in ("ilrl_env.cuh", "ilrl_chain.cuh", "ilrl_physics.cuh", "ilrl_constants.h",
                                                        "ilrl_model_data.h")] + [
    os.path.join(os.path.dirname(HERE), "include", "ilrl.h")]
# -ftz / -prec-div=false / -prec-sqrt=false: flush denormals, 2-ulp division and square root without their slow-path
# subroutines (+4..5 % throughput).  NOT --use_fast_math: sincosf / expf / atan2f keep full accuracy (the reward and
# observation parity bar is 1e-5 relative; the whole GPU suite passes with these flags).
# --register-usage-level=10 (ptxas spends more effort on keeping register use down): same results, +1.0..1.2 % at 65536
# envs, no change at 4096 (profiles/r2_v7_ptxas_flags_ab.txt; level 0: -2 %; --allow-expensive-optimizations: same code).
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-ftz=true",
              "-prec-div=false", "-prec-sqrt=false", "-Xptxas", "--register-usage-level=10", "-Xcompiler", "-fPIC", "-shared"]


def stale():
    return (not os.path.exists(SO)) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in DEPS)


def build(force=False, verbose=False):
    if not force and not stale():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, SRC, SRC_POLICY]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
