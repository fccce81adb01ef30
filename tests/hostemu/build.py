"""TEST INFRASTRUCTURE ONLY: build the host-emulation variant of the C-ABI library
(uno_api.cpp + plan.cpp + tests/hostemu/backend_host.cpp, plain g++, no CUDA)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "uno_b200", "csrc")
OUT = os.path.join(HERE, "libuno_hostemu.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, "uno_api.cpp"), os.path.join(CSRC, "plan.cpp"), os.path.join(CSRC, "config.cpp"), os.path.join(HERE, "backend_host.cpp")]
    deps = srcs + [os.path.join(CSRC, "backend.h"), os.path.join(CSRC, "plan.h"), os.path.join(CSRC, "config.h"), os.path.join(ROOT, "include", "uno_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", OUT] + srcs
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
