"""Op-level parity (run with `-m gpu`): every ggml op the two reference programs put into a graph, built through the
C ABI of include/ggml/ggml.h exactly the way main.cpp / rnn_text_generation.cpp call it, one op (or one short idiom of the
reference) per graph, compared with a numpy restatement of upstream ggml's semantics.  ggml ne = (ne0, ne1, ne2, ne3)
corresponds to a C-contiguous numpy array of shape (ne3, ne2, ne1, ne0).

These run the per-node plan (EXACT / EXACT_F32); the fused FAST plan is covered end to end in test_gpu_parity.py."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F32, F16, I32 = 0, 1, 26  # enum ggml_type (include/ggml/ggml.h)
vp = ctypes.c_void_p


class InitParams(ctypes.Structure):
    _fields_ = [("mem_size", ctypes.c_size_t), ("mem_buffer", vp), ("no_alloc", ctypes.c_bool)]


@pytest.fixture(scope="module")
def L():
    import ggml_experiments_b200 as G
    L = G.lib_ggml()
    assert L.ggml_b200_device_count() > 0, "no CUDA device: GPU tests need the B200 box"
    i64 = ctypes.c_int64
    sigs = {
        "ggml_init": (vp, [InitParams]), "ggml_free": (None, [vp]),
        "ggml_new_tensor_4d": (vp, [vp, ctypes.c_int, i64, i64, i64, i64]),
        "ggml_new_tensor_1d": (vp, [vp, ctypes.c_int, i64]),
        "ggml_new_f32": (vp, [vp, ctypes.c_float]),
        "ggml_get_data": (vp, [vp]), "ggml_nbytes": (ctypes.c_size_t, [vp]), "ggml_nelements": (i64, [vp]),
        "ggml_new_graph": (vp, [vp]), "ggml_build_forward_expand": (None, [vp, vp]),
        "ggml_graph_compute_with_ctx": (None, [vp, vp, ctypes.c_int]), "ggml_graph_release_plan": (None, [vp]),
        "ggml_norm": (vp, [vp, vp, ctypes.c_float]),
        "ggml_cont_4d": (vp, [vp, vp, i64, i64, i64, i64]),
        "ggml_reshape_2d": (vp, [vp, vp, i64, i64]), "ggml_reshape_3d": (vp, [vp, vp, i64, i64, i64]),
        "ggml_reshape_4d": (vp, [vp, vp, i64, i64, i64, i64]),
        "ggml_permute": (vp, [vp, vp] + [ctypes.c_int] * 4),
        "ggml_view_2d": (vp, [vp, vp, i64, i64, ctypes.c_size_t, ctypes.c_size_t]),
        "ggml_conv_2d": (vp, [vp, vp, vp] + [ctypes.c_int] * 6),
        "ggml_conv_depthwise_2d": (vp, [vp, vp, vp] + [ctypes.c_int] * 6),
        "ggml_b200_set_mode": (None, [ctypes.c_int]), "ggml_b200_get_mode": (ctypes.c_int, []),
    }
    for name in ("add", "sub", "mul", "div", "mul_mat", "repeat", "concat", "get_rows"):
        sigs["ggml_" + name] = (vp, [vp, vp, vp])
    for name in ("sqrt", "silu", "tanh", "soft_max", "cont", "transpose", "argmax"):
        sigs["ggml_" + name] = (vp, [vp, vp])
    sigs["ggml_b200_pool_mean_hw"] = (vp, [vp, vp])
    for name, (res, args) in sigs.items():
        getattr(L, name).restype = res
        getattr(L, name).argtypes = args
    return L


class Graph:
    """One ggml context + graph; leafs are filled from numpy arrays, the result is read back from the output's host data."""

    def __init__(self, L, mode):
        self.L, self.mode = L, mode
        self.prev = L.ggml_b200_get_mode()
        L.ggml_b200_set_mode(mode)
        self.ctx = L.ggml_init(InitParams(256 << 20, None, False))

    def leaf(self, a, dtype=F32):
        a = np.ascontiguousarray(a)
        ne = list(a.shape[::-1]) + [1] * (4 - a.ndim)
        t = self.L.ggml_new_tensor_4d(self.ctx, dtype, *ne)
        raw = {F32: np.float32, F16: np.float16, I32: np.int32}[dtype]
        buf = a.astype(raw)
        assert self.L.ggml_nbytes(t) == buf.nbytes
        ctypes.memmove(self.L.ggml_get_data(t), buf.ctypes.data, buf.nbytes)
        return t

    def run(self, out, shape, dtype=np.float32):
        gf = self.L.ggml_new_graph(self.ctx)
        self.L.ggml_build_forward_expand(gf, out)
        self.L.ggml_graph_compute_with_ctx(self.ctx, gf, 1)
        n = int(np.prod(shape))
        assert self.L.ggml_nelements(out) == n, (self.L.ggml_nelements(out), shape)
        res = np.ctypeslib.as_array(ctypes.cast(self.L.ggml_get_data(out), ctypes.POINTER(ctypes.c_float if dtype == np.float32 else ctypes.c_int32)),
                                    shape=(n,)).astype(dtype).reshape(shape).copy()
        self.L.ggml_graph_release_plan(gf)
        return res

    def close(self):
        self.L.ggml_free(self.ctx)
        self.L.ggml_b200_set_mode(self.prev)


@pytest.fixture(params=[1, 2], ids=["exact", "exact_f32"])
def g(L, request):
    gr = Graph(L, request.param)
    yield gr
    gr.close()


def rnd(seed, *shape):
    return np.random.default_rng(seed).normal(size=shape).astype(np.float32)


def close(got, ref, tol=2e-5):
    ref = np.asarray(ref, np.float64)
    err = np.abs(got.astype(np.float64) - ref).max()
    assert err <= tol * max(1.0, np.abs(ref).max()), err


# ---- elementwise, with the broadcast shapes the reference uses (main.cpp:810 bias (1,1,C,1); :1111 LN weight (C,1,1,1);
# rnn.cpp:209 bias (3U,1)) ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("op,fn", [("add", np.add), ("sub", np.subtract), ("mul", np.multiply), ("div", np.divide)])
@pytest.mark.parametrize("bshape", [(3, 5, 7, 9), (1, 5, 1, 1), (1, 1, 1, 9), (3, 1, 1, 1), (1, 1, 1, 1)])
def test_binary_broadcast(g, op, fn, bshape):
    a = rnd(1, 3, 5, 7, 9)
    b = rnd(2, *bshape) + (3.0 if op == "div" else 0.0)
    out = getattr(g.L, "ggml_" + op)(g.ctx, g.leaf(a), g.leaf(b))
    close(g.run(out, a.shape), fn(a.astype(np.float64), b.astype(np.float64)))


@pytest.mark.parametrize("op,fn", [("sqrt", np.sqrt), ("silu", lambda x: x / (1 + np.exp(-x))), ("tanh", np.tanh)])
def test_unary(g, op, fn):
    a = rnd(3, 2, 3, 17, 33) * 3
    if op == "sqrt":
        a = np.abs(a) + 1e-3
    close(g.run(getattr(g.L, "ggml_" + op)(g.ctx, g.leaf(a)), a.shape), fn(a.astype(np.float64)))


def test_reference_sigmoid_idiom(g):
    """rnn.cpp:51-55 writes sigmoid(x) as silu(x) / x."""
    a = rnd(4, 1, 1, 64, 96) * 4
    x = g.leaf(a)
    close(g.run(g.L.ggml_div(g.ctx, g.L.ggml_silu(g.ctx, x), x), a.shape), 1 / (1 + np.exp(-a.astype(np.float64))))


@pytest.mark.parametrize("n0", [1, 7, 64, 144, 1000])
def test_soft_max_and_norm_over_ne0(g, n0):
    a = rnd(5, 2, 3, 5, n0) * 4
    a64 = a.astype(np.float64)
    e = np.exp(a64 - a64.max(-1, keepdims=True))
    close(g.run(g.L.ggml_soft_max(g.ctx, g.leaf(a)), a.shape), e / e.sum(-1, keepdims=True))
    eps = 1e-5
    ref = (a64 - a64.mean(-1, keepdims=True)) / np.sqrt(a64.var(-1, keepdims=True) + eps)
    close(g.run(g.L.ggml_norm(g.ctx, g.leaf(a), eps), a.shape), ref, 1e-4)


def test_layernorm_idiom(g):
    """main.cpp:1006-1018: norm, then mul by repeat(weight), then add repeat(bias)."""
    a, w, b = rnd(6, 2, 4, 16, 96), rnd(7, 96), rnd(8, 96)
    x = g.L.ggml_norm(g.ctx, g.leaf(a), 1e-5)
    x = g.L.ggml_add(g.ctx, g.L.ggml_mul(g.ctx, x, g.leaf(w)), g.leaf(b))
    a64 = a.astype(np.float64)
    ref = (a64 - a64.mean(-1, keepdims=True)) / np.sqrt(a64.var(-1, keepdims=True) + 1e-5) * w + b
    close(g.run(x, a.shape), ref, 1e-4)


# ---- mul_mat: a [K,M,b2,b3] broadcast over b's dims 2,3 (main.cpp:1022 weights; :1075,1082 q.k^T and attn.v per head) ------
@pytest.mark.parametrize("K,M,N,a23,b23", [(64, 96, 50, (1, 1), (1, 1)), (144, 432, 256, (1, 1), (4, 3)), (36, 16, 16, (4, 3), (4, 3)),
                                           (5, 3, 2, (1, 1), (2, 1)), (256, 7, 129, (1, 1), (1, 2))])
@pytest.mark.parametrize("a_type", [F32, F16])
def test_mul_mat(g, K, M, N, a23, b23, a_type):
    a = rnd(9, a23[1], a23[0], M, K) / np.sqrt(K)
    b = rnd(10, b23[1], b23[0], N, K)
    out = g.L.ggml_mul_mat(g.ctx, g.leaf(a, a_type), g.leaf(b))
    a_eff = a.astype(np.float16).astype(np.float64) if a_type == F16 else a.astype(np.float64)
    # upstream rounds b to f16 when a is f16 (vec_dot_f16); EXACT_F32 deliberately does not
    b_eff = b.astype(np.float16).astype(np.float64) if (a_type == F16 and g.mode == 1) else b.astype(np.float64)
    ref = np.matmul(b_eff, np.swapaxes(np.broadcast_to(a_eff, b23[::-1] + (M, K)), -1, -2))
    close(g.run(out, b23[::-1] + (N, M)), ref, 2e-5)


# ---- data movement ----------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("axes", [(0, 1, 2, 3), (1, 0, 2, 3), (2, 0, 1, 3), (0, 2, 1, 3), (1, 2, 0, 3), (0, 1, 3, 2), (2, 1, 3, 0)])
def test_cont_of_permute(g, axes):
    """ggml_permute: result.ne[axes[i]] = a.ne[i] (main.cpp:723,745,758,1030,1047,1064,1091)."""
    a = rnd(11, 2, 3, 5, 8)
    out = g.L.ggml_cont(g.ctx, g.L.ggml_permute(g.ctx, g.leaf(a), *axes))
    # numpy axis j (from the slow end) holds ggml dim 3-j; ggml result dim axes[i] takes source dim i
    src_of = [0] * 4
    for i, ax in enumerate(axes):
        src_of[ax] = i
    perm = [3 - src_of[3 - j] for j in range(4)]
    ref = a.transpose(perm)
    np.testing.assert_array_equal(g.run(out, ref.shape), ref)


def test_transpose_reshape_cont_4d(g):
    a = rnd(12, 2, 3, 6, 8)
    out = g.L.ggml_cont_4d(g.ctx, g.L.ggml_transpose(g.ctx, g.leaf(a)), 12, 4, 3, 2)
    ref = a.transpose(0, 1, 3, 2).reshape(2, 3, 4, 12)
    np.testing.assert_array_equal(g.run(out, ref.shape), ref)
    r3 = g.L.ggml_reshape_3d(g.ctx, g.leaf(a), 16, 9, 2)
    out = g.L.ggml_add(g.ctx, r3, g.L.ggml_reshape_3d(g.ctx, g.leaf(a), 16, 9, 2))
    np.testing.assert_array_equal(g.run(out, (2, 9, 16)), (a + a).reshape(2, 9, 16))


def test_unfold_fold_idiom(g):
    """main.cpp:721-768 folding <-> unfolding with 2x2 patches, generalised to a batch: the round trip is the identity."""
    N, C, H, W, p = 2, 8, 6, 4, 2
    a = rnd(13, N, C, H, W)
    Lg, x = g.L, g.leaf(a)
    nh, nw = H // p, W // p
    # unfold (main.cpp:721-747): (W,H,C,N) -> (C, nh*nw, p*p, N)
    t = Lg.ggml_reshape_4d(g.ctx, x, p, nw, p, nh * C * N)
    t = Lg.ggml_cont(g.ctx, Lg.ggml_permute(g.ctx, t, 0, 2, 1, 3))
    t = Lg.ggml_reshape_4d(g.ctx, t, p * p, nh * nw, C, N)
    patches = Lg.ggml_cont(g.ctx, Lg.ggml_permute(g.ctx, t, 2, 1, 0, 3))
    ref = a.reshape(N, C, nh, p, nw, p).transpose(0, 3, 5, 2, 4, 1).reshape(N, p * p, nh * nw, C)
    np.testing.assert_array_equal(g.run(patches, ref.shape), ref)
    # fold (main.cpp:749-768) is the inverse
    t = Lg.ggml_cont(g.ctx, Lg.ggml_permute(g.ctx, g.leaf(ref), 2, 1, 0, 3))
    t = Lg.ggml_reshape_4d(g.ctx, t, p, p, nw, nh * C * N)
    t = Lg.ggml_cont(g.ctx, Lg.ggml_permute(g.ctx, t, 0, 2, 1, 3))
    back = Lg.ggml_reshape_4d(g.ctx, t, W, H, C, N)
    np.testing.assert_array_equal(g.run(back, a.shape), a)


def test_repeat_concat_get_rows_view_argmax(g):
    Lg = g.L
    a = rnd(14, 2, 3, 5, 8)
    one = Lg.ggml_new_f32(g.ctx, 1.0)
    np.testing.assert_array_equal(g.run(Lg.ggml_repeat(g.ctx, one, g.leaf(a)), a.shape), np.ones_like(a))  # rnn.cpp:244
    row = rnd(15, 1, 3, 1, 8)
    np.testing.assert_array_equal(g.run(Lg.ggml_repeat(g.ctx, g.leaf(row), g.leaf(a)), a.shape), np.broadcast_to(row, a.shape))
    b = rnd(16, 2, 4, 5, 8)
    np.testing.assert_array_equal(g.run(Lg.ggml_concat(g.ctx, g.leaf(a), g.leaf(b)), (2, 7, 5, 8)), np.concatenate([a, b], 1))  # main.cpp:1219
    table, ids = rnd(17, 65, 256), np.array([3, 0, 64, 64, 17], np.int32)
    np.testing.assert_array_equal(g.run(Lg.ggml_get_rows(g.ctx, g.leaf(table), g.leaf(ids, I32)), (5, 256)), table[ids])  # rnn.cpp:200
    m = rnd(18, 6, 24)
    v = Lg.ggml_view_2d(g.ctx, g.leaf(m), 8, 6, 24 * 4, 8 * 4)
    np.testing.assert_array_equal(g.run(Lg.ggml_cont(g.ctx, v), (6, 8)), m[:, 8:16])
    logits = rnd(19, 37, 65)
    np.testing.assert_array_equal(g.run(Lg.ggml_argmax(g.ctx, g.leaf(logits)), (37,), np.int32), logits.argmax(-1))  # rnn.cpp:74-77


# ---- convolutions (main.cpp:788 conv_2d, :798 conv_depthwise_2d; kernels F16, activations rounded to F16 by upstream im2col) -
def conv_ref(x, w, s, p, groups=1):
    N, C, H, W = x.shape
    OC, IC, KH, KW = w.shape
    OH, OW = (H + 2 * p - KH) // s + 1, (W + 2 * p - KW) // s + 1
    xp = np.zeros((N, C, H + 2 * p, W + 2 * p))
    xp[:, :, p:p + H, p:p + W] = x
    out = np.zeros((N, OC, OH, OW))
    for kh in range(KH):
        for kw in range(KW):
            win = xp[:, :, kh:kh + s * OH:s, kw:kw + s * OW:s]
            if groups == 1:
                out += np.einsum("nchw,oc->nohw", win, w[:, :, kh, kw])
            else:
                out += win * w[None, :, 0, kh, kw, None, None]
    return out


@pytest.mark.parametrize("N,C,OC,H,W,k,s,p", [(2, 3, 16, 32, 32, 3, 2, 1), (1, 16, 64, 16, 16, 1, 1, 0), (3, 24, 24, 9, 11, 3, 1, 1),
                                              (1, 8, 8, 7, 5, 3, 2, 1), (2, 160, 32, 8, 8, 1, 1, 0)])
def test_conv_2d(g, N, C, OC, H, W, k, s, p):
    x, w = rnd(20, N, C, H, W), (rnd(21, OC, C, k, k) / np.sqrt(C * k * k)).astype(np.float16)
    out = g.L.ggml_conv_2d(g.ctx, g.leaf(w, F16), g.leaf(x), s, s, p, p, 1, 1)
    x_eff = x.astype(np.float16) if g.mode == 1 else x
    ref = conv_ref(x_eff.astype(np.float64), w.astype(np.float64), s, p)
    close(g.run(out, ref.shape), ref, 2e-5)


@pytest.mark.parametrize("N,C,H,W,s", [(2, 16, 16, 16, 1), (1, 64, 32, 32, 2), (3, 8, 9, 7, 1), (2, 8, 9, 7, 2)])
def test_conv_depthwise_2d(g, N, C, H, W, s):
    x, w = rnd(22, N, C, H, W), (rnd(23, C, 1, 3, 3) / 3).astype(np.float16)
    out = g.L.ggml_conv_depthwise_2d(g.ctx, g.leaf(w, F16), g.leaf(x), s, s, 1, 1, 1, 1)
    x_eff = x.astype(np.float16) if g.mode == 1 else x
    ref = conv_ref(x_eff.astype(np.float64), w.astype(np.float64), s, 1, groups=C)
    close(g.run(out, ref.shape), ref, 2e-5)


def test_pool_mean_hw(g):
    a = rnd(24, 3, 40, 8, 8)
    close(g.run(g.L.ggml_b200_pool_mean_hw(g.ctx, g.leaf(a)), (3, 40, 1, 1)), a.astype(np.float64).mean((2, 3), keepdims=True))


def test_same_graph_twice_with_new_inputs(L):
    """Leafs of the compute context are re-uploaded on every compute (main.cpp:627-640 refills the image and recomputes)."""
    gr = Graph(L, 1)
    try:
        a = rnd(25, 4, 32)
        x = gr.leaf(a)
        out = L.ggml_silu(gr.ctx, x)
        gf = L.ggml_new_graph(gr.ctx)
        L.ggml_build_forward_expand(gf, out)
        for k in range(3):
            cur = (a * (k + 1)).astype(np.float32)
            ctypes.memmove(L.ggml_get_data(x), cur.ctypes.data, cur.nbytes)
            L.ggml_graph_compute_with_ctx(gr.ctx, gf, 1)
            got = np.ctypeslib.as_array(ctypes.cast(L.ggml_get_data(out), ctypes.POINTER(ctypes.c_float)), shape=(a.size,)).reshape(a.shape)
            close(got, cur.astype(np.float64) / (1 + np.exp(-cur.astype(np.float64))))
        L.ggml_graph_release_plan(gf)
    finally:
        gr.close()
