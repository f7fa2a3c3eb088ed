"""Host-side logic and the C-ABI surface, no GPU: the library loads, exports every symbol that
include/lpgnn.h declares, and refuses to compute without a device (no CPU fallback)."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "lpgnn.h")).read()
    return sorted(set(re.findall(r"LPGNN_API\s+[\w\s\*]+?(lpgnn_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 12
    assert sorted(_lib.SIGNATURES.keys()) == syms                   # ctypes table == header
    lib = _lib.load()
    for s in syms:
        assert hasattr(lib, s), s
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (lpgnn_\w+)", out))
    assert exported == set(syms)                                    # nothing undeclared leaks out either
    assert lib.lpgnn_version() == 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_no_cpu_fallback_without_device():
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import _lib, ops
    from lpgnn_b200.graph import BipartiteCSR
    lib = _lib.load()
    assert lib.lpgnn_device_info(None, None, None) == -3            # LPGNN_ENODEVICE
    assert "no CPU fallback" in _lib.last_error()
    x = torch.zeros(4, 8)
    ptr = torch.zeros(5, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.spmm((ptr, ptr, x, 4), x)
    g = BipartiteCSR.from_edge_index(torch.tensor([[0, 1], [1, 0]]), torch.tensor([1.0, 2.0]), (2, 2))
    assert not g.is_cuda and g.nnz() == 2 and g.t().sparse_sizes() == (2, 2)
    with pytest.raises(RuntimeError, match="host COO"):
        g.views()


def test_sass_contains_blackwell_tensor_and_tma_instructions():
    """The node transforms must be tcgen05/TMA kernels, not mma.sync (B200_PROFILING.md mnemonics).  The ONE place that
    may carry the register-fragment HMMA is the 16-bit input layer (csrc/conv_in_mma.cu, its two kernels): a reduction of length 16 is a
    single MMA step whose accumulators have to reach registers for the store anyway -- the TMEM drain was the measured
    bottleneck of its tcgen05 form; see the file header and DESIGN.md."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import _lib
    exe = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([exe, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    per_function = re.split(r"\n\s*Function : ", sass)
    hmma_in = [f.split("\n", 1)[0] for f in per_function[1:] if "HMMA." in f.replace("UTCHMMA", "")]
    assert hmma_in and all("conv_in_mma_kernel" in name or "conv_in_mma_regb_kernel" in name for name in hmma_in), hmma_in
    tc = [f.split("\n", 1)[0] for f in per_function[1:] if "UTCHMMA" in f]
    assert any("gemm_tc_kernel" in n for n in tc) and any("gemm_x2_kernel" in n for n in tc)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "lp-gnn_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f


def test_synthetic_lp_layout_and_invariants():
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import synth
    lp = synth.config_lp("C1")
    assert lp.c_feas.shape == (1000, 8) and lp.v_feas.shape == (2000, 8)
    assert lp.c_feas.dtype == np.float32 and 9000 < lp.nnz <= 10_500
    assert np.abs(lp.a_data).max() <= 1 and np.abs(lp.c_feas).max() <= 1       # dataset.py:235-238
    key = lp.row * lp.n + lp.col
    assert (np.diff(key) > 0).all()                                            # row-major, no duplicates
    for feas, y in ((lp.c_feas, lp.y_s), (lp.v_feas, lp.y_t)):                 # dataset.py:203-207
        assert set(np.unique(feas[:, 5])) <= {-1.0, 0.0, 1.0}
        assert (y[feas[:, 5] != 0] != 0).all() and (y[feas[:, 7] != 0] != 2).all()
    again = synth.config_lp("C1")
    assert np.array_equal(again.row, lp.row) and np.array_equal(again.v_feas, lp.v_feas)   # seeded
    pop = synth.lp_population(50)
    assert len(pop) == 50 and all(n == 2 * m for m, n, _, _ in pop)


def test_features_against_golden():
    import scipy.sparse as sp
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import features
    gold = os.path.join(ROOT, "tests", "golden")
    for name in ("tiny_5x7", "small_300x600", "bounds_40x90"):
        z = np.load(os.path.join(gold, f"lp_features_{name}.npz"))
        m, n = z["in_shape"]
        A = sp.csr_matrix((z["in_A_data"], z["in_A_indices"], z["in_A_indptr"]), shape=(m, n))
        c, b_l, A2, b_u, l, u = features.scale_lp(z["in_c"], z["in_b_l"], A, z["in_b_u"], z["in_l"], z["in_u"])
        np.testing.assert_array_equal(A2.data, z["out_A_data"])
        np.testing.assert_array_equal(c, z["out_c"])
        v, cf = features.node_features(c, b_l, A2, b_u, l, u)
        np.testing.assert_allclose(v, z["out_v_feas"], rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(cf, z["out_c_feas"], rtol=1e-12, atol=1e-14)
        np.testing.assert_array_equal(v[:, [5, 7]], z["out_v_feas"][:, [5, 7]])


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's
    keys, the oracle port timed on the host cores, no GPU needed."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "C1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300, env=env, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "LPs/s" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "LPs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None
    # a non-zero rank under torchrun exits 0 without work or output
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "C1",
                          "--steps", "1"], capture_output=True, text=True, timeout=120, env=dict(env, RANK="1", WORLD_SIZE="2"),
                         cwd=root)
    assert out.returncode == 0 and not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a box without a GPU")
def test_bench_gpu_arm_fails_loudly_without_a_device():
    """No CPU fallback: the product arm of bench.py must not print a number on a box without a GPU."""
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-cpu"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert "CUDA" in out.stderr or "cuda" in out.stderr


def test_ctypes_bindings_match_the_header_signatures():
    """Every entry point of include/lpgnn.h is bound in _lib.py with explicit argtypes whose count equals the C
    parameter count (a call through ctypes without argtypes would truncate 64-bit pointers and sizes), and every
    `size_t` / `uint64_t` / `const char*` return type is declared."""
    import ctypes

    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import _lib
    lib = _lib.load()
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "lpgnn.h")).read(), flags=re.S)
    seen = 0
    for m in re.finditer(r"LPGNN_API\s+([\w\s\*]+?)(lpgnn_\w+)\s*\((.*?)\)\s*;", src, flags=re.S):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        n = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
        fn = getattr(lib, name)
        assert fn.argtypes is not None and len(fn.argtypes) == n, (name, n, fn.argtypes)
        if ret == "size_t":
            assert fn.restype is ctypes.c_size_t, name
        elif ret == "uint64_t":
            assert fn.restype is ctypes.c_uint64, name
        elif ret.replace(" ", "") == "constchar*":
            assert fn.restype is ctypes.c_char_p, name
        seen += 1
    assert seen == len(_header_symbols())


def test_integration_md_stubs_match_the_header():
    """The ctypes stubs a maintainer would copy from INTEGRATION.md declare as many arguments as the header does."""
    import ctypes as C
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "lpgnn.h")).read(), flags=re.S)
    hdr = {}
    for m in re.finditer(r"LPGNN_API\s+[\w\s\*]+?(lpgnn_\w+)\s*\((.*?)\)\s*;", src, flags=re.S):
        params = m.group(2).strip()
        hdr[m.group(1)] = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    ns = {"C": C, "_p": C.c_void_p, "_i32": C.c_int32, "_i64": C.c_int64}
    stubs = re.findall(r"_lib\.(lpgnn_\w+)\.argtypes = (.*)", doc)
    assert len(stubs) >= 6
    for name, expr in stubs:
        assert len(eval(expr, ns)) == hdr[name], name                      # noqa: S307 (repo-owned document)
    for name in re.findall(r"`(lpgnn_[a-z0-9_]+)`", doc):
        assert any(h == name or h.startswith(name) for h in hdr), f"INTEGRATION.md names an unknown entry point {name}"
