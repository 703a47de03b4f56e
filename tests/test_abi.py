"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly the symbols include/wd_b200.h declares.
No compute calls are made here (no GPU in the build container)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "wd_b200.h")


@pytest.fixture(scope="module")
def lib_path():
    from worddiffusion_b200.build import build_library
    return build_library()


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wd_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_bound_and_exported(lib_path):
    from worddiffusion_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    assert sorted(_lib.SIGNATURES) == declared, "ctypes binding and header disagree"
    l = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(l, name), f"{name} not exported by libwd_b200.so"


def test_library_has_no_torch_or_cudart_dependency(lib_path):
    out = subprocess.run(["ldd", lib_path], capture_output=True, text=True).stdout
    names = " ".join(line.split()[0] for line in out.splitlines() if line.strip())  # library names only (load addresses are random hex)
    assert "torch" not in names and "libcudart" not in names and "c10" not in names


def test_version_and_error_string(lib_path):
    from worddiffusion_b200 import _lib
    l = _lib.lib()
    assert l.wd_version() >= 1
    assert l.wd_op_gemm_block_n() % 16 == 0
    # invalid arguments are rejected before any CUDA call
    rc = l.wd_engine_create(None, None)
    assert rc == -1 and b"null" in l.wd_last_error()


def test_sass_is_blackwell_native(lib_path):
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG (B200_PROFILING.md, 'What proves a Blackwell-native kernel')."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", lib_path], capture_output=True, text=True).stdout
    for mnem in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnem in sass, mnem
    assert "sm_100a" in subprocess.run([cuobjdump, "-lelf", lib_path], capture_output=True, text=True).stdout


def test_no_cpu_fallback():
    import torch
    from worddiffusion_b200 import _lib
    from worddiffusion_b200.unet import UNetModel, default_args
    m = UNetModel(image_size=(64, 256), in_channels=4, model_channels=320, out_channels=4, num_res_blocks=1,
                  attention_resolutions=(1, 1), channel_mult=(1, 1), num_heads=4, num_classes=339, context_dim=320,
                  vocab_size=53, args=default_args("cpu"), max_seq_len=10)
    x = torch.zeros(1, 4, 8, 32)
    with pytest.raises(_lib.WdError):
        m(x, None, timesteps=torch.tensor([1]), context=torch.zeros(1, 10, dtype=torch.long), y=torch.tensor([0]))
    with pytest.raises(RuntimeError):
        m.input_blocks[1][0](x)  # parameter holders carry no arithmetic


def test_product_package_never_touches_the_oracle_or_the_reference():
    """The oracle is test infrastructure: nothing under worddiffusion_b200/ (Python or CUDA) may import, open or name oracle/ or
    /root/reference; bench.py may execute the oracle only in its cpu_baseline / --impl reference legs."""
    pkg = os.path.join(ROOT, "worddiffusion_b200")
    pat = re.compile(r"^\s*(from|import)\s+\S*(oracle|ref_shims|weights)\b|/root/reference|sys\.path")
    hits = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                for i, line in enumerate(open(os.path.join(d, f), errors="replace"), 1):
                    if pat.search(line):
                        hits.append(f"{f}:{i}: {line.strip()[:100]}")
    assert not hits, hits
    bench = open(os.path.join(ROOT, "bench.py")).read()
    # bench.py imports the oracle's UNet in ONE place (_load_oracle), reached only from the CPU-baseline function (also the
    # `--impl reference` arm) and from the labelled torch-eager comparator -- never from the measured product path
    assert bench.count("import unet_oracle") == 1
    assert bench.split("import unet_oracle")[0].rsplit("\ndef ", 1)[1].startswith("_load_oracle(")
    callers = set()
    for chunk in bench.split("\ndef ")[1:]:
        name, body = chunk.split("(", 1)
        if "_load_oracle()" in body:
            callers.add(name)
    assert callers == {"cpu_oracle_throughput", "gpu_eager_reference"}, callers
