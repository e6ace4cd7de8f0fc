"""The registry plumbing of an 8th backend: tools/gen_backend_ph.py re-derives the reference's perfect hashes
(lib/nnc/cmd/build-cmd.rb:303-386, lib/nnc/cmd/ccv_nnc_cmd.inc:152-190) and the stand-alone host dispatches through
init_map[_ccv_nnc_cmd_ph(cmd)].backends[_ccv_nnc_cmd_backend_ph(backend)] exactly as lib/nnc/ccv_nnc_cmd.c does."""
import os
import subprocess
import sys

import pytest

from ccv_b200 import abi, nnc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_BACKENDS = {"CPU_OPT": 0x46deb194, "CPU_REF": 0x3d9883e5, "GPU_CUBLAS": 0x9b8cfed, "GPU_CUDNN": 0x854b679a, "GPU_NCCL": 0x7afed9c7, "GPU_REF": 0x5f19790a, "MPS": 0xb2f325e2}


def test_generator_reproduces_the_reference_backend_hash_and_extends_it():
    if not os.path.exists("/root/reference/lib/nnc/cmd/ccv_nnc_cmd.inc"):
        pytest.skip("the reference tree is only present in the build container")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_backend_ph.py"), "--check-only"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0, out.stdout
    assert "7-backend hash reproduced: (backend >> 15) % 7 + 0" in out.stdout
    assert "8-backend hash" in out.stdout


def test_host_dispatch_goes_through_the_generated_tables():
    nnc.init()
    ok = nnc.lib().ccv_nnc_cmd_ok
    # every command the header lists is registered under GPU_SM100 and under no other backend id
    for name in dir(abi):
        if name.startswith("CCV_NNC_") and (name.endswith("_FORWARD") or name.endswith("_BACKWARD")) and isinstance(getattr(abi, name), int):
            cmd = getattr(abi, name)
            if ok(cmd, abi.CCV_NNC_BACKEND_GPU_SM100):
                for other in REFERENCE_BACKENDS.values():
                    assert ok(cmd, other) == 0, (name, hex(other))
    assert ok(abi.CCV_NNC_CONVOLUTION_FORWARD, abi.CCV_NNC_BACKEND_GPU_SM100) == 1
    assert ok(abi.CCV_NNC_CONVOLUTION_FORWARD, 0x12345678) == 0          # not a backend id
    assert ok(0x0badc0de, abi.CCV_NNC_BACKEND_GPU_SM100) == 0           # not a command id
    # generated fragments are committed next to the integration notes
    for f in ("ccv_nnc_backend.h", "ccv_nnc_cmd_backend.inc", "ccv_nnc_cmd_sm100_init.inc"):
        text = open(os.path.join(ROOT, "integration", f)).read()
        assert "CCV_NNC_BACKEND_GPU_SM100" in text
    assert "CCV_NNC_BACKEND_COUNT = 8" in open(os.path.join(ROOT, "integration", "ccv_nnc_backend.h")).read()


def test_patched_cmd_inc_reslots_every_registration_call(tmp_path):
    """integration/patch_cmd_inc.py (what build-cmd.rb would regenerate for an 8th backend): every `_register_command_X_backend_Y` call of the
    reference's generated file lands on the slot the 8-slot hash gives Y, the command slot is untouched, backend_init_map is ordered by
    slot (lib/nnc/ccv_nnc_cmd.c:61-66 requires backend_init_map[ph(b)].backend == b) and every SM100 registration of the header is present."""
    import re
    src = "/root/reference/lib/nnc/cmd/ccv_nnc_cmd.inc"
    if not os.path.exists(src):
        pytest.skip("the reference tree is only present in the build container")
    dst = str(tmp_path / "ccv_nnc_cmd.inc")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "integration", "patch_cmd_inc.py"), src, dst], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert out.returncode == 0, out.stdout
    text, ref = open(dst).read(), open(src).read()
    backends = re.findall(r'\{\.name = "(CCV_NNC_BACKEND_[A-Z0-9_]+)", \.backend = (0x[0-9a-f]+)\}', text)
    assert len(backends) == 8

    def ph8(b):  # integration/ccv_nnc_cmd_backend.inc
        return ((b >> 2) % 8) if (b % 2 == 0) else ((b >> 9) % 8)
    for slot, (name, ident) in enumerate(backends):
        assert ph8(int(ident, 16)) == slot, (name, slot)
    slot_of = {name: i for i, (name, _) in enumerate(backends)}
    call = re.compile(r"_register_command_(CCV_NNC_[A-Z0-9_]+?)_backend_(CCV_NNC_BACKEND_[A-Z0-9_]+)\(&\(init_map\[(\d+)\]\.backends\[(\d+)\]\)\);")
    new_calls, old_calls = call.findall(text), call.findall(ref)
    n_sm100 = len(re.findall(r"X\(CCV_NNC_[A-Z0-9_]+\)", open(os.path.join(ROOT, "include", "ccv_nnc_sm100.h")).read()))  # CCV_NNC_SM100_COMMANDS
    assert len(old_calls) == 334 and len(new_calls) == 334 + n_sm100 and n_sm100 >= 94
    for cmd, backend, i, j in new_calls:
        assert int(j) == slot_of[backend], (cmd, backend, j)
    # the command slot of every pre-existing call is the reference's
    assert sorted((c, b, i) for c, b, i, _ in old_calls) == sorted((c, b, i) for c, b, i, _ in new_calls if b != "CCV_NNC_BACKEND_GPU_SM100")
    # an SM100 call uses the command slot its command already has
    cmd_slot = {c: i for c, _, i, _ in old_calls}
    sm100 = [(c, i) for c, b, i, _ in new_calls if b == "CCV_NNC_BACKEND_GPU_SM100"]
    assert len(sm100) == n_sm100 and all(cmd_slot[c] == i for c, i in sm100)
