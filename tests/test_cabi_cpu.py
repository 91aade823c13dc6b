"""CPU-side checks of the boundary: the C-ABI library builds/loads and exports every symbol include/desmo_b200.h declares;
host logic (scheduler, shapes, error behaviour without a GPU).  No compute calls."""
import ctypes

import numpy as np
import pytest

from desmo_b200 import _lib
from oracle import desmo_oracle as orc


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _lib.declared_symbols()
    assert len(declared) >= 20
    assert set(declared) == set(_lib.SIGNATURES), (set(declared) ^ set(_lib.SIGNATURES))
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.desmo_version()


def test_library_term_counts_match_reference_facts():
    lib = _lib.load()
    for (r, p) in [(4, 3), (4, 2), (2, 2), (3, 4), (2, 7), (8, 1)]:
        assert lib.desmo_num_terms(r, p) == orc.number_of_terms(r, p)
        assert lib.desmo_padded_k(r, p) % 16 == 0 and lib.desmo_padded_k(r, p) >= orc.number_of_terms(r, p) + 3 * r
    assert lib.desmo_num_terms(8, 2) == 45 and lib.desmo_padded_k(8, 2) == 80  # BASELINE's "8 modes": K = 69
    # BASELINE's larger libraries: "8 modes" with p = 3, "32 modes", the sweep's r = 64 (GEMM path)
    assert lib.desmo_num_terms(8, 3) == 165 and lib.desmo_num_terms(32, 2) == 561 and lib.desmo_num_terms(64, 2) == 2145
    assert lib.desmo_padded_k(32, 2) == 672
    assert lib.desmo_num_terms(64, 3) < 0  # K = 48097 > DESMO_MAX_K: reported as unsupported, never silently truncated
    assert lib.desmo_num_terms(2, 8) < 0 and lib.desmo_num_terms(0, 2) < 0 and lib.desmo_num_terms(65, 1) < 0


def test_compile_time_monomial_tables_match_the_runtime_enumeration():
    """The register-resident chain-rule kernels (fused_fp32.cu: chain_rule_reg_kernel<R, P>) fold POOL_DATA's term order
    (CYL:376-434) into the instruction stream; every such table must equal the run-time enumeration the other kernels use."""
    lib = _lib.load()
    assert lib.desmo_selftest_tables() == 6


@pytest.mark.parametrize("r,p", [(4, 2), (2, 2), (2, 3), (2, 4), (3, 2), (3, 3)])
def test_unrolled_chain_rule_sweep_matches_the_library_derivative(r, p):
    """The reverse sweep the specialised chain-rule kernels run per point (chain_sweep_ct, the same inlined code on host and device)
    against d/dPhi of POOL_DATA's monomials (CYL:376-434) in float64: dPhi_i = sum_j D_j * d(prod_q Phi_idx_j[q]) / dPhi_i."""
    import itertools

    lib = _lib.load()
    combos = [()] + [c for d in range(1, p + 1) for c in itertools.combinations_with_replacement(range(r), d)]
    T = len(combos)
    assert T == orc.number_of_terms(r, p)
    rng = np.random.default_rng(r * 10 + p)
    for _ in range(20):
        D = rng.standard_normal(T).astype(np.float32)
        Phi = (rng.standard_normal(r) * 1.5).astype(np.float32)
        want = np.zeros(r)
        for j, c in enumerate(combos):
            for i in set(c):
                k = c.count(i)
                rest = np.prod([float(Phi[q]) for q in c if q != i]) if any(q != i for q in c) else 1.0
                want[i] += float(D[j]) * k * float(Phi[i]) ** (k - 1) * rest
        got = np.zeros(r, np.float32)
        fp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))  # noqa: E731
        assert lib.desmo_selftest_chain_sweep(r, p, fp(D), fp(Phi), fp(got)) == 0
        assert np.allclose(got, want, rtol=2e-5, atol=2e-5 * np.abs(want).max()), (got, want)
    assert lib.desmo_selftest_chain_sweep(5, 2, fp(D), fp(Phi), fp(got)) != 0  # no specialised kernel: reported, not guessed


def test_shape_validation_and_error_strings():
    lib = _lib.load()
    bad = _lib.make_shape(1000, 100, 4, 2, ld=1000)  # pitch not a multiple of 128
    assert lib.desmo_red_count(ctypes.byref(bad)) == -1
    assert b"ld=" in lib.desmo_last_error()
    ok = _lib.make_shape(1000, 100, 4, 2)
    K, Kp = 27, 32
    assert lib.desmo_red_count(ctypes.byref(ok)) == Kp * ok.mld + 1 + 16 + 12
    assert ok.ld % 128 == 0 and ok.mld % 16 == 0
    # dispatch: fused tcgen05 kernel for K <= 32, m <= 1024; GEMM path beyond; explicit paths are refused where they do not apply
    sel = lambda *a, **k: lib.desmo_selected_path(ctypes.byref(_lib.make_shape(*a, **k)))  # noqa: E731
    assert sel(1000, 100, 4, 2) == _lib.PATH_TC and sel(1000, 100, 4, 3) == _lib.PATH_GEMM and sel(1000, 2000, 4, 2) == _lib.PATH_GEMM
    assert sel(1000, 100, 32, 2) == _lib.PATH_GEMM and sel(1000, 100, 4, 3, path=_lib.PATH_FP32) == _lib.PATH_FP32
    assert sel(1000, 100, 4, 3, path=_lib.PATH_TC) < 0 and sel(1000, 100, 32, 2, path=_lib.PATH_FP32) < 0


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import desmo_b200

    with pytest.raises(desmo_b200.DesmoError):
        desmo_b200.DESMO(100, 10, 2, 2)
    lib = _lib.load()
    n = ctypes.c_size_t()
    assert lib.desmo_workspace_bytes(ctypes.byref(_lib.make_shape(1000, 100, 4, 2)), ctypes.byref(n)) == 3  # DESMO_ERR_CUDA
    out = ctypes.c_void_p()
    assert lib.desmo_session_create(1000, 100, 4, 2, 0, 0, ctypes.byref(out)) == 3


def test_plateau_scheduler_matches_torch():
    import torch

    from desmo_b200.trainer import PlateauScheduler

    prm = [torch.nn.Parameter(torch.zeros(1)) for _ in range(4)]
    opt = torch.optim.Adamax([{"params": [p], "lr": lr} for p, lr in zip(prm, (1e-2, 1e-3, 1e-2, 1e3))])
    ref = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", patience=2, factor=0.1, min_lr=1e-6)
    mine = PlateauScheduler((1e-2, 1e-3, 1e-2, 1e3), patience=2)
    rng = np.random.default_rng(0)
    metric = 10.0
    for it in range(200):
        metric = metric * (0.9 if it % 17 == 0 else 1.0) + 1e-6 * rng.standard_normal()
        ref.step(metric)
        mine.step(metric)
        assert np.allclose(mine.lrs, [g["lr"] for g in opt.param_groups], rtol=1e-12), it


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of desmo_shape / desmo_plateau / desmo_peer have the header's sizes and field offsets (checked by compiling the
    header with the C compiler): a silent ABI drift between include/desmo_b200.h and the Python binding would corrupt launches."""
    import os
    import shutil
    import subprocess

    from desmo_b200.engine import _PeerDesc
    from desmo_b200.trainer import _PlateauState

    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "desmo_b200.h"\n'
                   'int main(void) {\n'
                   '  printf("shape %zu %zu %zu %zu %zu\\n", sizeof(desmo_shape), offsetof(desmo_shape, n_global), offsetof(desmo_shape, m), offsetof(desmo_shape, nF), offsetof(desmo_shape, path));\n'
                   '  printf("plateau %zu %zu %zu %zu %zu %zu\\n", sizeof(desmo_plateau), offsetof(desmo_plateau, lrs), offsetof(desmo_plateau, threshold), offsetof(desmo_plateau, num_bad), offsetof(desmo_plateau, every), offsetof(desmo_plateau, reductions));\n'
                   '  printf("peer %zu %zu %zu %zu\\n", sizeof(desmo_peer), offsetof(desmo_peer, red_ptrs), offsetof(desmo_peer, flag_ptrs), offsetof(desmo_peer, state));\n'
                   '  return 0;\n}\n')
    exe = tmp_path / "layout"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run([cc, "-std=c99", "-I", inc, str(src), "-o", str(exe)], check=True)
    out = dict((ln.split()[0], [int(v) for v in ln.split()[1:]]) for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    S, P, Q = _lib.Shape, _PlateauState, _PeerDesc
    assert out["shape"] == [ctypes.sizeof(S), S.n_global.offset, S.m.offset, S.nF.offset, S.path.offset]
    assert out["plateau"] == [ctypes.sizeof(P), P.lrs.offset, P.threshold.offset, P.num_bad.offset, P.every.offset, P.reductions.offset]
    assert out["peer"] == [ctypes.sizeof(Q), Q.red_ptrs.offset, Q.flag_ptrs.offset, Q.state.offset]
