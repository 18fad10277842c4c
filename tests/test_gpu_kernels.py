"""GPU: every L0 kernel of the C ABI against NumPy on the same seeded inputs."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from sypha_b200 import _lib  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def ptr(t):
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


@pytest.mark.parametrize("n", [1, 31, 257, 100_003, 1_300_000])
def test_elementwise_and_ratio_test(lib, n):
    """elem_min_mult_dev / corrector_rhs_dev / alpha_max_dev (sypha_solver_utils.h:7-24); n up to
    1.3M exercises the range where the reference's alpha_max_dev silently bails (n > 262144)."""
    r = np.random.default_rng(n)
    x, s = r.uniform(0.1, 5, n), r.uniform(0.1, 5, n)
    dx, ds = r.normal(size=n), r.normal(size=n)
    X, S, DX, DS = dev(x), dev(s), dev(dx), dev(ds)
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    assert lib.sb200_k_elem_min_mult(ptr(X), ptr(S), ptr(out), n, stream()) == 0
    assert np.array_equal(out.cpu().numpy(), -x * s)
    assert lib.sb200_k_corrector_rhs(ptr(DX), ptr(DS), 0.37, 2.5, ptr(out), n, stream()) == 0
    # nvcc contracts -dx*ds + sigma*mu into one FMA (so does the reference's kernel): 1 ulp of the product
    assert np.all(np.abs(out.cpu().numpy() - (-dx * ds + 0.37 * 2.5)) <= 2.3e-16 * (1.0 + np.abs(dx * ds)))
    res = torch.zeros(2, dtype=torch.float64, device="cuda")
    h = np.zeros(2)
    assert lib.sb200_k_alpha_max(ptr(X), ptr(DX), ptr(S), ptr(DS), n, ptr(res), h.ctypes.data, stream()) == 0
    ep = np.min(-x[dx < 0] / dx[dx < 0]) if (dx < 0).any() else np.finfo(float).max
    ed = np.min(-s[ds < 0] / ds[ds < 0]) if (ds < 0).any() else np.finfo(float).max
    assert h[0] == ep and h[1] == ed          # min is exact: bit-equal
    assert np.array_equal(res.cpu().numpy(), h)


def test_ratio_test_empty_set(lib):
    n = 1000
    x = np.ones(n)
    dx = np.abs(np.random.default_rng(0).normal(size=n))      # no negative direction
    X, DX = dev(x), dev(dx)
    res = torch.zeros(2, dtype=torch.float64, device="cuda")
    h = np.zeros(2)
    lib.sb200_k_alpha_max(ptr(X), ptr(DX), ptr(X), ptr(DX), n, ptr(res), h.ctypes.data, stream())
    assert h[0] == np.finfo(float).max and h[1] == np.finfo(float).max     # DBL_MAX, utils.cu:77-78


@pytest.mark.parametrize("shape", [(7, 13, 0.5), (200, 1200, 0.02), (1000, 11000, 0.05), (64, 5000, 0.3)])
def test_spmv_csr_csc_jacobi(lib, shape):
    m, n, dens = shape
    r = np.random.default_rng(m)
    A = sp.random(m, n, density=dens, format="csr", random_state=r, data_rvs=lambda k: r.integers(-3, 4, k).astype(float))
    A.eliminate_zeros()
    A.sort_indices()
    At = A.tocsc()
    At.sort_indices()
    x, y0 = r.normal(size=n), r.normal(size=m)
    v = r.normal(size=m)
    offs, inds, vals = dev(A.indptr.astype(np.int32)), dev(A.indices.astype(np.int32)), dev(A.data)
    X, Y = dev(x), dev(y0)
    assert lib.sb200_k_spmv_csr(m, ptr(offs), ptr(inds), ptr(vals), ptr(X), ptr(Y), -1.5, 0.5, stream()) == 0
    ref = -1.5 * (A @ x) + 0.5 * y0
    assert np.allclose(Y.cpu().numpy(), ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    cp, rows, cvals = dev(At.indptr.astype(np.int32)), dev(At.indices.astype(np.int32)), dev(At.data)
    V, Z = dev(v), dev(x.copy())
    assert lib.sb200_k_spmv_csc(n, ptr(cp), ptr(rows), ptr(cvals), ptr(V), ptr(Z), 2.0, -1.0, stream()) == 0
    ref = 2.0 * (A.T @ v) - x
    assert np.allclose(Z.cpu().numpy(), ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    d = r.uniform(0.1, 10, n)
    D, diag = dev(d), torch.empty(m, dtype=torch.float64, device="cuda")
    assert lib.sb200_k_jacobi_diag(m, ptr(offs), ptr(inds), ptr(vals), ptr(D), ptr(diag), stream()) == 0
    ref = (A.multiply(A) @ d)
    assert np.allclose(diag.cpu().numpy(), ref, rtol=1e-12)


def _spd(n, seed):
    r = np.random.default_rng(seed)
    B = r.normal(size=(n, n + 8))
    return B @ B.T + n * 1e-3 * np.eye(n)


@pytest.mark.parametrize("n", [5, 64, 65, 200, 1000, 2111])
def test_potrf_potrs(lib, n):
    """dense Cholesky + solve (replaces cusolverDnDgetrf/Dgetrs, dense_linear.cpp:179,195)."""
    ld = (n + 63) // 64 * 64
    M = _spd(n, n)
    P = np.eye(ld)
    P[:n, :n] = M
    A = dev(P)
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert lib.sb200_k_potrf(n, ptr(A), ld, ptr(info), stream()) == 0
    torch.cuda.synchronize()
    assert int(info.item()) == 0
    Lg = np.tril(A.cpu().numpy())[:n, :n]
    Lr = np.linalg.cholesky(M)
    assert np.max(np.abs(Lg - Lr)) <= 1e-10 * np.abs(Lr).max()
    b = np.random.default_rng(1).normal(size=n)
    bp = np.zeros(ld)
    bp[:n] = b
    Bv = dev(bp)
    assert lib.sb200_k_potrs(n, ptr(A), ld, ptr(Bv), stream()) == 0
    torch.cuda.synchronize()
    xs = Bv.cpu().numpy()
    ref = np.linalg.solve(M, b)
    assert np.max(np.abs(xs[:n] - ref)) <= 1e-9 * np.abs(ref).max()
    assert np.all(xs[n:] == 0.0)
    # repeated solves reuse the factor (epoch flags must re-arm)
    for k in range(3):
        b2 = np.random.default_rng(10 + k).normal(size=n)
        bp[:n] = b2
        Bv = dev(bp)
        lib.sb200_k_potrs(n, ptr(A), ld, ptr(Bv), stream())
        torch.cuda.synchronize()
        ref = np.linalg.solve(M, b2)
        assert np.max(np.abs(Bv.cpu().numpy()[:n] - ref)) <= 1e-9 * np.abs(ref).max()


def test_potrf_reports_non_positive_pivot(lib):
    n, ld = 130, 192
    M = _spd(n, 3)
    M[100, 100] = -1.0
    P = np.eye(ld)
    P[:n, :n] = M
    A = dev(P)
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    lib.sb200_k_potrf(n, ptr(A), ld, ptr(info), stream())
    torch.cuda.synchronize()
    assert int(info.item()) == 101          # 1-based index of the first bad pivot


@pytest.mark.parametrize("shape", [(64, 32), (200, 1200), (500, 1000), (130, 77)])
def test_syrk_dmma(lib, shape):
    """FP64 tensor-core SYRK  C = A diag(d) A'  against NumPy."""
    m, k = shape
    ld, lda = (m + 63) // 64 * 64, (k + 31) // 32 * 32
    r = np.random.default_rng(m + k)
    A = np.zeros((ld, lda))
    A[:m, :k] = r.normal(size=(m, k))
    d = np.zeros(lda)
    d[:k] = r.uniform(0.01, 100, k)
    Ad, Dd = dev(A), dev(d)
    Cd = torch.zeros((ld, ld), dtype=torch.float64, device="cuda")
    assert lib.sb200_k_syrk(m, k, ptr(Ad), lda, ptr(Dd), ptr(Cd), ld, stream()) == 0
    torch.cuda.synchronize()
    ref = (A * d) @ A.T
    got = Cd.cpu().numpy()
    il = np.tril_indices(ld)
    assert np.max(np.abs(got[il] - ref[il])) <= 1e-12 * np.abs(ref).max()


@pytest.mark.parametrize("shape", [(40, 300, 0.1), (700, 40000, 0.01), (30000, 70000, 0.0005)])
def test_blocked_pattern_products(lib, shape):
    """+/-1 matrices loaded for the PCG strategy use the pattern-only, shared-memory-staged products
    (sb200_blocked.cu); they must agree with the value-carrying kernels / SciPy on A x and A' v, for one
    and for several blocks of the dense vector (n > 16384, m > 13312)."""
    import sypha_b200 as sb
    m, n0, dens = shape
    r = np.random.default_rng(m)
    A0 = sp.random(m, n0, density=dens, format="csr", random_state=r, data_rvs=lambda k: np.ones(k))
    A = sp.hstack([A0, -sp.identity(m)], format="csr")       # standard form [A0 | -I]
    A.sort_indices()
    n = n0 + m
    env = sb.SyphaEnvironment(linearSolverStrategy="pcg")
    node = sb.SyphaNodeSparse.from_csr(m, n, n0, A.indptr.astype(np.int32), A.indices.astype(np.int32),
                                       A.data.astype(np.float64), np.ones(n), np.ones(m), env)
    ws = sb.IpmWorkspace()
    sb.initializeIpmWorkspace(ws)
    try:
        node.copyModelOnDevice(ws)
        info = (C.c_longlong * 16)()
        lib.sb200_model_info(ws.handle, info, 16)
        assert info[12] == 1, "blocked pattern was not built for a +/-1 PCG model"
        if n > 16384:
            assert info[13] > 1
        if m > 13312:
            assert info[14] > 1
        x, v = r.normal(size=n), r.normal(size=m)
        X, Y = dev(x), torch.empty(m, dtype=torch.float64, device="cuda")
        assert lib.sb200_ws_spmv(ws.handle, 0, ptr(X), ptr(Y)) == 0
        ref = A @ x
        assert np.allclose(Y.cpu().numpy(), ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
        V, Z = dev(v), torch.empty(n, dtype=torch.float64, device="cuda")
        assert lib.sb200_ws_spmv(ws.handle, 1, ptr(V), ptr(Z)) == 0
        ref = A.T @ v
        assert np.allclose(Z.cpu().numpy(), ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
        # deterministic: bit-identical on repetition
        Z2 = torch.empty(n, dtype=torch.float64, device="cuda")
        lib.sb200_ws_spmv(ws.handle, 1, ptr(V), ptr(Z2))
        assert torch.equal(Z, Z2)
    finally:
        sb.releaseIpmWorkspace(ws)


def test_general_coefficients_keep_value_kernels(lib):
    """a cut row with a coefficient 2 must not take the pattern-only path."""
    import sypha_b200 as sb
    m, n = 30, 200
    r = np.random.default_rng(5)
    A = sp.random(m, n, density=0.2, format="csr", random_state=r, data_rvs=lambda k: np.ones(k))
    A = A.tolil()
    A[3, 7] = 2.0
    A = A.tocsr()
    A.sort_indices()
    env = sb.SyphaEnvironment(linearSolverStrategy="pcg")
    node = sb.SyphaNodeSparse.from_csr(m, n, n, A.indptr.astype(np.int32), A.indices.astype(np.int32),
                                       A.data.astype(np.float64), np.ones(n), np.ones(m), env)
    ws = sb.IpmWorkspace()
    sb.initializeIpmWorkspace(ws)
    try:
        node.copyModelOnDevice(ws)
        info = (C.c_longlong * 16)()
        lib.sb200_model_info(ws.handle, info, 16)
        assert info[12] == 0
        x = r.normal(size=n)
        X, Y = dev(x), torch.empty(m, dtype=torch.float64, device="cuda")
        assert lib.sb200_ws_spmv(ws.handle, 0, ptr(X), ptr(Y)) == 0
        assert np.allclose(Y.cpu().numpy(), A @ x, rtol=1e-12)
    finally:
        sb.releaseIpmWorkspace(ws)
