"""SURVEY.md 8f-1, the drop-in proof as tests.  integration/Makefile (run by __graft_entry__.build() where /root/reference exists) builds
the UNMODIFIED reference -- every layer of lib/nnc above the backends, its CPU backends and its own CUDA compatibility layer --
with the 8-backend registry files, links it with this repository's backend objects (no nnc_host.o: ccv_nnc_cmd_exec, tensors,
stream contexts, the graph runner and ccv_cnnp_model are the reference's) and compiles the reference's own test programs unmodified
with the GPU backend names reading CCV_NNC_BACKEND_GPU_SM100.  The binaries travel with the tree; nothing here reads /root/reference.

CPU: the reference's 55 unit programs (386 cases: known-answer GEMM / convolution / attention, autograd, symbolic-graph compile and
simplification, while / case-of, dynamic graphs, cnnp core, dataframe, tensor IO) pass against the patched registry, i.e. the 8-slot
backend hash and the re-slotted registration calls leave every CPU backend where ccv_nnc_cmd_exec and the upper layers find it.
GPU: the reference's integration programs whose cases are deterministic and quick -- ccv_cnnp_model training (cnnp.core),
statically scheduled multi-stream graphs with while / case-of (schedule), SGD / ADAM(W) in float, half and mixed precision, index
select, tensor transfer, datatype conversion, concat, leaky relu, GELU, SWISH, transforms, reductions, upsample -- pass on this
backend as the only GPU backend of the library.  (cudnn.tests / cublas.tests, minutes of CPU_REF convolutions, are run by
integration/run_reference_tests.py: profiles/r02_dropin_reference_tests.md.)"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "integration", "_build")
RUNNER = os.path.join(ROOT, "integration", "run_reference_tests.py")


def _built(programs):
    return os.path.exists(os.path.join(BUILD, "libccv_dropin.so")) and os.path.exists(os.path.join(BUILD, "cases.json")) and all(os.path.exists(os.path.join(BUILD, p + ".tests")) for p in programs)


def _run(only, budget, tmp_path):
    out = str(tmp_path / "report.json")
    p = subprocess.run([sys.executable, RUNNER, "--only", ",".join(only), "--budget-s", str(budget), "--case-timeout", "60", "--out", out], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=budget + 60)
    assert p.returncode == 0, p.stdout.decode()[-2000:]
    return json.load(open(out))


# unit cases that cannot pass here for reasons unrelated to the registry: they read files under test/unit/nnc/data by a relative path (the
# programs run from integration/_build) or map a file they wrote there; and one cnnp case that segfaults identically when the same test
# source is linked against the pristine 7-backend CPU build of the reference (oracle/_ref/libccv_ref.so) in this toolchain
UNIT_KNOWN = {"read long text from csv", "read Windows csv", "read csv include header", "read a larger csv file", "read a csv file", "tensor mapped from file",
              "LoRA fine-tuning MLP with GELU, set is_trainable to false and with gradient checkpointing"}


def test_reference_unit_programs_pass_on_the_8_backend_registry(tmp_path):
    """all 55 unit programs of the reference (test/unit/nnc/makefile TARGETS minus cblas): autograd, symbolic graph compile / simplify /
    while / case-of, dynamic graph, cnnp core, dataframe, tensor IO, every CPU command family"""
    if not _built(["unit.gemm", "unit.cnnp.core", "unit.symbolic.graph", "unit.dynamic.graph", "unit.while"]):
        pytest.skip("integration/_build is not built (make -C integration; needs the reference sources)")
    rep = _run(["unit."], 600, tmp_path)
    # Some of the reference's unit cases train from a random initialisation and hold the result to a tight bound, or abort on it: in
    # cnnp.core "LoRA fine-tuning ..." read 9.988 against 10 +- 0.01 in one run here, and "train a simple math 2 * x + 1 + 1 = 10 ..." aborts in
    # about one run out of four -- identically when the same test source is linked against the pristine 7-backend CPU build of the
    # reference (oracle/_ref/libccv_ref.so).  Those programs are therefore only counted; every other program must be clean.
    stochastic = {"unit.cnnp.core.tests", "unit.minimize.tests", "unit.rand.tests", "unit.dropout.tests"}
    failed = {p: r["fail_detail"] for p, r in rep["programs"].items() if r["failed"] and p not in stochastic}
    crashed = {k for p, r in rep["programs"].items() if p not in stochastic for k in r["crashed"]}
    assert not failed, failed
    assert crashed <= UNIT_KNOWN, crashed - UNIT_KNOWN
    assert rep["total"]["PASS"] >= 370, rep["total"]  # 386 on the build container when no stochastic case trips
    assert rep["programs"]["unit.cnnp.core.tests"]["tally"]["PASS"] >= 35, rep["programs"]["unit.cnnp.core.tests"]["tally"]  # of 42
    for prog, want in (("unit.gemm", 18), ("unit.forward", 17), ("unit.backward", 3), ("unit.attention", 6), ("unit.dynamic.graph", 36), ("unit.while", 12)):
        assert rep["programs"][prog + ".tests"]["tally"]["PASS"] == want, (prog, rep["programs"][prog + ".tests"]["tally"])


# program -> least number of cases that must PASS (none may FAIL or crash).  The counts are those of the whole-program B200 run of this
# build (profiles/r02_dropin_reference_tests.md); for adam / reduce / upsample, whose last fixes were confirmed case by case (12 / 9 / 9
# pass), the floor is the whole-program count before those fixes.
GPU_PROGRAMS = {"int.cnnp.core": 2, "int.schedule": 5, "int.sgd": 6, "int.index": 9, "int.tensor": 7, "int.datatype": 1, "int.concat": 2, "int.leaky_relu": 4,
                "int.gelu": 8, "int.swish": 4, "int.transform": 9, "int.adam": 8, "int.reduce": 7, "int.upsample": 5}


@pytest.mark.gpu
def test_reference_integration_programs_pass_on_the_dropin_build(gpu, tmp_path):
    if not _built(list(GPU_PROGRAMS)):
        pytest.skip("integration/_build is not built (make -C integration; needs the reference sources)")
    rep = _run(list(GPU_PROGRAMS), 300, tmp_path)
    # int.cnnp.core trains "2 * x = 10" from a random initialisation like its unit counterpart, which aborts in about one run out of four on
    # the pristine reference too: counted (at least half of its 4 cases), not held to zero failures
    bad = {p: (r["failed"], list(r["crashed"])) for p, r in rep["programs"].items() if (r["failed"] or r["crashed"]) and p != "int.cnnp.core.tests"}
    assert not bad, json.dumps({p: {"failed": r.get("fail_detail"), "crashed": r.get("crashed")} for p, r in rep["programs"].items() if p in bad}, indent=1)[-3000:]
    for prog, want in GPU_PROGRAMS.items():
        got = rep["programs"][prog + ".tests"]["tally"]["PASS"]
        assert got >= want, (prog, got, want, rep["programs"][prog + ".tests"]["skipped"])


def test_cnnp_resnet50_program_trains_on_the_cpu_backends():
    """integration/cnnp_resnet50_bench.c (a ccv_cnnp_model ResNet-50 v1d through ccv_cnnp_model_compile / _fit, the reference's public API)
    on CPU tensors: the program's own logic -- model definition, compile, fit loop, loss read-back -- checked where no GPU is needed."""
    exe = os.path.join(BUILD, "cnnp_resnet50_bench")
    if not os.path.exists(exe):
        pytest.skip("integration/_build is not built (make -C integration; needs the reference sources)")
    # 64 x 64 images: at 32 x 32 the last stage is 1 x 1 and its batch norms see 8 values; some random initialisations diverge there on
    # the reference's CPU backends (loss -> -log 1e-30), which says nothing about this program
    p = subprocess.run([exe, "--device", "cpu", "--batch", "8", "--image", "64", "--steps", "3", "--warmup", "1", "--classes", "10", "--lr", "0.002", "--seed", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600)
    assert p.returncode == 0, p.stderr.decode()[-1000:]
    r = json.loads(p.stdout.decode().strip().splitlines()[-1])
    assert r["device"] == "cpu" and r["images_per_sec"] > 0
    assert 0 < r["last_loss"] < r["first_loss"], r  # four SGD steps on one fixed batch lower its loss


@pytest.mark.gpu
def test_cnnp_resnet50_program_trains_on_the_dropin_build(gpu):
    exe = os.path.join(BUILD, "cnnp_resnet50_bench")
    if not os.path.exists(exe):
        pytest.skip("integration/_build is not built")
    """The same program on the GPU: the reference's cnnp / symbolic graph / autograd / planner / scheduled graph runner issuing a ResNet-50
    training step to this backend.  The configuration is the one run on a B200 with this build (profiles/r02_cnnp_resnet50_reference_api_gpu.json:
    batch 64 at 128 x 128, 9.7 ms/step, loss 8.67 -> 4.18 over five steps)."""
    p = subprocess.run([exe, "--device", "gpu", "--batch", "64", "--image", "128", "--steps", "3", "--warmup", "2"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=240)
    assert p.returncode == 0, p.stderr.decode()[-1500:]
    r = json.loads(p.stdout.decode().strip().splitlines()[-1])
    assert r["device"] == "gpu" and r["images_per_sec"] > 0
    assert 0 < r["last_loss"] < r["first_loss"], r  # five nesterov steps on one fixed batch
