"""The contraction planner (jues.jl_b200/csrc/contract_plan.h) is pure host logic: build its C++ check with
g++ and run it on the CPU -- every contraction string of the coupled-cluster sweep, plain and batched."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_contract_plans_against_brute_force(tmp_path):
    exe = str(tmp_path / "contract_plan_test")
    src = os.path.join(ROOT, "tests", "host", "contract_plan_test.cpp")
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe, src], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failures" in r.stdout
