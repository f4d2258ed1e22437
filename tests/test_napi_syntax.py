"""The N-API shim cannot run here (no node); check that it compiles against the hand-declared
N-API subset and binds only symbols the C ABI header declares."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAPI = os.path.join(ROOT, "you_can_not_recommend_b200", "napi")


def test_shim_syntax():
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-Wall", "-DYCNR_NAPI_MIN", "-I", NAPI,
                        "-I", os.path.join(ROOT, "include"), os.path.join(NAPI, "ycnr_napi.cc")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_shim_uses_only_declared_abi():
    header = open(os.path.join(ROOT, "include", "ycnr_als.h")).read()
    declared = set(re.findall(r"\b(ycnr_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", header, flags=re.S)))
    used = set(re.findall(r"\b(ycnr_[a-z_0-9]+)\s*\(", open(os.path.join(NAPI, "ycnr_napi.cc")).read()))
    assert used and used <= declared, used - declared
    gyp = open(os.path.join(NAPI, "binding.gyp")).read()
    assert '"target_name": "cpp_utils"' in gyp            # same addon slot as upstream binding.gyp:4
