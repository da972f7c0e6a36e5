"""GPU: every C-ABI kernel against a plain torch fp32 statement of the same op (bf16 tolerances
stated per test).  Index-valued results are checked bit-exactly."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def cos(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()


@pytest.fixture(scope="module")
def ops():
    from m3l_b200 import ops as _ops
    return _ops


# --------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("bn", [0, 64, 128, 256])
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (1000, 520, 200), (2560, 768, 256), (1536, 256, 192)])
def test_gemm_kmajor(ops, M, N, K, bn):
    torch.manual_seed(0)
    a = torch.randn(M, K, device=DEV).bfloat16()
    b = torch.randn(N, K, device=DEV).bfloat16()
    out = ops.gemm(a, b, bn=bn)
    assert rel_err(out, a.float() @ b.float().T) < 1e-2  # bf16 output rounding


@pytest.mark.parametrize("bn,splits", [(64, 1), (128, 4), (256, 7), (0, 3)])
def test_gemm_wgrad_mnmajor_splitk(ops, bn, splits):
    torch.manual_seed(1)
    a = torch.randn(1000, 264, device=DEV).bfloat16()
    b = torch.randn(1000, 200, device=DEV).bfloat16()
    out = torch.ones(264, 200, device=DEV)
    ops.gemm(a, b, mn_major=True, out=out, accumulate=True, splits=splits, bn=bn)
    assert rel_err(out, 1 + a.float().T @ b.float()) < 1e-5  # fp32 accumulate of exact bf16 products


def test_gemm_epilogues(ops):
    torch.manual_seed(2)
    M, N, K = 777, 512, 256
    a = torch.randn(M, K, device=DEV).bfloat16()
    b = (torch.randn(N, K, device=DEV) * 0.05).bfloat16()
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV).bfloat16()
    z = a.float() @ b.float().T
    out = ops.gemm(a, b, bias=bias, residual=res)
    assert rel_err(out, z + bias + res.float()) < 1e-2
    twin = torch.zeros_like(out)
    out_b = ops.gemm(a, b, bias=bias, residual=res, out2=twin)       # second copy of the output
    assert torch.equal(out_b, out) and torch.equal(twin, out)
    dgelu = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    h = ops.gemm(a, b, bias=bias, act=ops.GELU_FWD, aux_out=dgelu)
    zz = (z + bias).requires_grad_(True)
    F.gelu(zz).sum().backward()
    assert rel_err(dgelu, zz.grad) < 1e-2 and rel_err(h, F.gelu(z + bias)) < 1e-2   # aux = GELU'(pre-activation)
    h2 = ops.gemm(a, b, bias=bias, act=ops.GELU_FWD)
    assert torch.equal(h2, h)
    cs = torch.ones(N, device=DEV)
    g = ops.gemm(a, b, act=ops.GELU_BWD, aux_in=dgelu, colsum_out=cs)
    assert rel_err(g, z * dgelu.float()) < 1e-2
    assert rel_err(cs, 1 + g.float().sum(0)) < 1e-4  # fused bias gradient = column sums of the bf16 output
    # fused row-dot: out = a @ b.T (bf16), dot[row, c] = sum_j out[row, 64c + j] * side[row, 64c + j]
    side = torch.randn(M, N, device=DEV).bfloat16()
    dot = torch.full((M, N // 64), 7.0, device=DEV)
    o2 = ops.gemm(a, b, dot_side=side, dot_out=dot)
    assert torch.equal(o2, ops.gemm(a, b))
    assert rel_err(dot, (o2.float() * side.float()).reshape(M, N // 64, 64).sum(-1)) < 1e-5
    big_a = torch.randn(40000, K, device=DEV).bfloat16()
    cs2 = torch.zeros(N, device=DEV)
    g2 = ops.gemm(big_a, b, colsum_out=cs2, bn=128)
    assert rel_err(cs2, g2.float().sum(0)) < 1e-4
    f32 = ops.gemm(a, b, bias=bias, out_dtype=torch.float32)
    assert rel_err(f32, z + bias) < 1e-5
    inplace = res.clone()
    ops.gemm(a, b, bias=bias, residual=inplace, out=inplace)
    assert rel_err(inplace, z + bias + res.float()) < 1e-2


@pytest.mark.parametrize("M,N,K", [(12288 + 77, 1024, 256), (49152, 1024, 256), (20000, 512, 192)])
def test_gemm_gelu_fwd_16_warp_kernel(ops, M, N, K):
    """Feed-forward forward shapes that take the dedicated 16-epilogue-warp GELU kernel (gemm_gelu.cu:
    weight-stationary, N % 256 == 0, K <= 256, >= 2 tiles per SM), ragged M tail included: GELU and GELU'
    against torch, and elementwise against the generic epilogue (M3L_GELU16 is read once, so the generic
    path is reached through BN = 128 here)."""
    torch.manual_seed(3)
    a = torch.randn(M, K, device=DEV).bfloat16()
    b = (torch.randn(N, K, device=DEV) * 0.07).bfloat16()
    bias = torch.randn(N, device=DEV)
    dg = torch.full((M, N), 9.0, device=DEV, dtype=torch.bfloat16)
    h = ops.gemm(a, b, bias=bias, act=ops.GELU_FWD, aux_out=dg, bn=256)
    z = a.float() @ b.float().T + bias
    zz = z.clone().requires_grad_(True)
    F.gelu(zz).sum().backward()
    assert rel_err(h, F.gelu(z)) < 1e-2 and rel_err(dg, zz.grad) < 1e-2
    dg2 = torch.empty_like(dg)
    h2 = ops.gemm(a, b, bias=bias, act=ops.GELU_FWD, aux_out=dg2, bn=128)
    # same fp32 accumulators and the same A&S erf; only the packed-vs-scalar association of the last products
    # differs, i.e. at most one bf16 ulp on a few elements
    assert (h.float() - h2.float()).abs().max() <= 2 ** -7 * max(1.0, float(h2.float().abs().max()))
    assert (dg.float() - dg2.float()).abs().max() <= 2 ** -7 * 2.0
    assert float((h != h2).float().mean()) < 0.02 and float((dg != dg2).float().mean()) < 0.02
    # backward through the GELU at the same shape (generic epilogue; a 16-warp variant with a single staging
    # buffer per warp measured no faster: 49 vs 49 us): dPre = (dY W2) * GELU', fused column sums (bias gradient)
    dy = torch.randn(M, K, device=DEV).bfloat16()
    cs = torch.ones(N, device=DEV)
    dpre = ops.gemm(dy, b, act=ops.GELU_BWD, aux_in=dg, colsum_out=cs, bn=256)
    zb = dy.float() @ b.float().T
    assert rel_err(dpre, zb * dg.float()) < 1e-2
    assert rel_err(cs, 1 + dpre.float().sum(0)) < 1e-4
    cs128 = torch.ones(N, device=DEV)
    dpre128 = ops.gemm(dy, b, act=ops.GELU_BWD, aux_in=dg, colsum_out=cs128, bn=128)
    assert torch.equal(dpre, dpre128) and rel_err(cs, cs128) < 1e-5
    assert torch.equal(ops.gemm(dy, b, act=ops.GELU_BWD, aux_in=dg, bn=256), dpre)


# ------------------------------------------------------------------------------- mask indices
# The weight-stationary variants the benchmark shape (B = 256 per GPU) selects: gemm_bf16_kernel<256,0,0,0,1> (decoder
# QKV, M = 49152), <256,0,0,3,1> / <128,0,0,3,1> (to_pixels / to_tactiles, fp32 epilogue, M = 15360 / 31232).
@pytest.mark.parametrize("M,N,K,f32", [(49152, 768, 256, False), (15360, 768, 256, True), (31232, 192, 256, True),
                                       (49152, 256, 256, False)])
def test_gemm_weight_stationary_variants_at_bench_shape(ops, M, N, K, f32):
    torch.manual_seed(11)
    a = torch.randn(M, K, device=DEV).bfloat16()
    b = (torch.randn(N, K, device=DEV) * 0.06).bfloat16()
    bias = torch.randn(N, device=DEV)
    out = ops.gemm(a, b, bias=bias, out_dtype=torch.float32 if f32 else torch.bfloat16)
    ref = a.float() @ b.float().T + bias
    assert rel_err(out, ref) < (1e-5 if f32 else 1e-2)
    # every 128-row tile is right, not only the largest element
    blk = (out.float() - ref).abs().reshape(M // 128, -1).amax(1)
    assert blk.max().item() < (1e-4 if f32 else 6e-2) * ref.abs().max().item()


def _ln_mlp_reference(x, gamma, beta, w1, b1, w2, b2):
    """fp32 statement of x + W2 GELU(W1 LN(x) + b1) + b2 with the kernel's bf16 rounding points."""
    xf = x.float()
    mean = xf.mean(1, keepdim=True)
    var = ((xf - mean) ** 2).mean(1, keepdim=True)
    rstd = torch.rsqrt(var + 1e-5)
    xn = ((xf - mean) * rstd * gamma + beta).bfloat16()
    pre = (xn.float() @ w1.float().T + b1).requires_grad_(True)
    hf = F.gelu(pre)
    hf.sum().backward()
    h = hf.detach().bfloat16()
    out = xf + h.float() @ w2.float().T + b2
    return out, torch.cat([mean, rstd], 1), xn, h, pre.grad


@pytest.mark.parametrize("M,hidden", [(128, 128), (128, 1024), (777, 512), (2560, 512), (2560, 1024), (49152, 1024),
                                      (148 * 128 * 2 + 5, 256)])
def test_ln_mlp_fused_block(ops, M, hidden):
    torch.manual_seed(5)
    D = 256
    x = (torch.randn(M, D, device=DEV) * 1.5 + 0.3).bfloat16()
    gamma = 1 + 0.2 * torch.randn(D, device=DEV)
    beta = 0.1 * torch.randn(D, device=DEV)
    w1 = (torch.randn(hidden, D, device=DEV) * 0.08).bfloat16()
    b1 = 0.5 * torch.randn(hidden, device=DEV)
    w2 = (torch.randn(D, hidden, device=DEV) * 0.05).bfloat16()
    b2 = 0.3 * torch.randn(D, device=DEV)
    ref_out, ref_stats, ref_xn, ref_h, ref_gp = _ln_mlp_reference(x, gamma, beta, w1, b1, w2, b2)
    out = ops.ln_mlp_fwd(x, gamma, beta, w1, b1, w2, b2)
    assert rel_err(out, ref_out) < 1e-2
    blk = (out.float() - ref_out).abs().amax(1)
    assert blk.max().item() < 6e-2 * ref_out.abs().max().item()
    out2, stats, xn, h, gp = ops.ln_mlp_fwd(x, gamma, beta, w1, b1, w2, b2, save=True)
    assert rel_err(out2, ref_out) < 1e-2
    assert rel_err(stats, ref_stats) < 1e-4
    assert rel_err(xn, ref_xn) < 1e-2 and rel_err(h, ref_h) < 1e-2 and rel_err(gp, ref_gp) < 1e-2
    # the two variants compute the same block
    assert rel_err(out2, out) < 1e-2
    # accumulate modes: in place on the residual stream, and into a buffer that already holds x (reduce-add epilogue)
    xi = x.clone()
    o3 = ops.ln_mlp_fwd(xi, gamma, beta, w1, b1, w2, b2, out=xi)
    assert o3.data_ptr() == xi.data_ptr() and rel_err(xi, ref_out) < 1e-2
    assert (xi.float() - ref_out).abs().amax(1).max().item() < 6e-2 * ref_out.abs().max().item()
    buf = x.clone()
    o4, stats4, xn4, h4, gp4 = ops.ln_mlp_fwd(x, gamma, beta, w1, b1, w2, b2, save=True, out=buf, out_has_x=True)
    assert rel_err(o4, ref_out) < 1e-2 and torch.equal(stats4, stats) and torch.equal(xn4, xn)
    assert torch.equal(h4, h) and torch.equal(gp4, gp)
    # inference form accumulating into a copy of x
    buf5 = x.clone()
    o5 = ops.ln_mlp_fwd(x, gamma, beta, w1, b1, w2, b2, out=buf5, out_has_x=True)
    assert rel_err(o5, ref_out) < 1e-2
    assert (o5.float() - ref_out).abs().amax(1).max().item() < 6e-2 * ref_out.abs().max().item()


def test_ln_mlp_rejects_unsupported_shapes(ops):
    from m3l_b200._lib import M3LError
    x = torch.zeros(128, 384, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(M3LError):
        ops.ln_mlp_fwd(x, torch.ones(384, device=DEV), torch.zeros(384, device=DEV),
                       torch.zeros(768, 384, device=DEV, dtype=torch.bfloat16), torch.zeros(768, device=DEV),
                       torch.zeros(384, 768, device=DEV, dtype=torch.bfloat16), torch.zeros(384, device=DEV))
    assert not ops.ln_mlp_supported(384, 768) and ops.ln_mlp_supported(256, 1024)


@pytest.mark.parametrize("B,segs", [(37, [(0, 64, 60), (64, 64, 61), (128, 64, 61)]), (5, [(0, 64, 60)]),
                                    (9, [(0, 25, 20), (25, 25, 20)]), (3, [(0, 300, 17)])])
def test_mask_indices_bit_exact(ops, B, segs):
    g = torch.Generator().manual_seed(3)
    n = sum(l for _, l, _ in segs)
    noise = torch.rand(B, n, generator=g)
    masked, unmasked, slots = ops.mask_indices(noise.to(DEV), segs)
    m_ref, u_ref = [], []
    for off, l, nm in segs:
        perm = noise[:, off:off + l].argsort(dim=-1, stable=True) + off
        m_ref.append(perm[:, :nm]); u_ref.append(perm[:, nm:])
    m_ref, u_ref = torch.cat(m_ref, 1), torch.cat(u_ref, 1)
    assert torch.equal(masked.cpu(), m_ref) and torch.equal(unmasked.cpu(), u_ref)
    s = slots.cpu()
    br = torch.arange(B)[:, None]
    assert torch.equal(s[br, u_ref], torch.arange(u_ref.shape[1], dtype=torch.int32).expand(B, -1))
    assert torch.equal(s[br, m_ref], -(1 + torch.arange(m_ref.shape[1], dtype=torch.int32)).expand(B, -1))


def test_mask_indices_ties_are_stable(ops):
    noise = torch.tensor([[0.5, 0.25, 0.5, 0.25, 0.0, 0.5, 1.0, 0.25]])
    masked, unmasked, _ = ops.mask_indices(noise.to(DEV), [(0, 8, 5)], want_slots=False)
    assert masked.cpu().tolist() == [[4, 1, 3, 7, 0]] and unmasked.cpu().tolist() == [[2, 5, 6]]


# ------------------------------------------------------------------------ patchify + LayerNorm
def _patchify(x, p1, p2):
    b, c, H, W = x.shape
    h, w = H // p1, W // p2
    return x.reshape(b, c, h, p1, w, p2).permute(0, 2, 4, 3, 5, 1).reshape(b, h * w, p1 * p2 * c)


def test_patch_layernorm_gather(ops):
    torch.manual_seed(4)
    B = 6
    t1, t2 = torch.rand(B, 12, 32, 32, device=DEV), torch.rand(B, 12, 32, 32, device=DEV)
    gamma, beta = torch.randn(192, device=DEV), torch.randn(192, device=DEV)
    ps = ops.make_patch_source([t1, t2], 4, 4, token_base=64)
    idx = torch.stack([torch.randperm(128)[:7] + 64 for _ in range(B)]).to(DEV)
    pad = torch.zeros(B, 3, dtype=torch.int64, device=DEV)
    idx_full = torch.cat([pad, idx], 1).contiguous()  # tactile columns start at col0 = 3
    out, xhat = ops.patch_layernorm(ps, B, 7, gamma, beta, tok_idx=idx_full, col0=3)
    patches = torch.cat([_patchify(t1, 4, 4), _patchify(t2, 4, 4)], 1)
    sel = patches[torch.arange(B, device=DEV)[:, None], idx - 64].reshape(B * 7, 192)
    assert rel_err(xhat, F.layer_norm(sel, (192,))) < 1e-2
    assert rel_err(out, F.layer_norm(sel, (192,), gamma, beta)) < 1e-2
    # all tokens (no index list), image geometry
    img = torch.rand(3, 12, 64, 64, device=DEV)
    g2, b2 = torch.randn(768, device=DEV), torch.randn(768, device=DEV)
    ps2 = ops.make_patch_source([img], 8, 8, token_base=0)
    out2, _ = ops.patch_layernorm(ps2, 3, 64, g2, b2, want_xhat=False)
    assert rel_err(out2, F.layer_norm(_patchify(img, 8, 8).reshape(-1, 768), (768,), g2, b2)) < 1e-2


# ----------------------------------------------------------------------------------- LayerNorm
@pytest.mark.parametrize("D,fp32_in", [(256, False), (128, False), (384, True), (1024, False)])
def test_layernorm_fwd_bwd(ops, D, fp32_in):
    torch.manual_seed(5)
    M = 1237
    x = torch.randn(M, D, device=DEV) * 2 + 0.5
    x = x if fp32_in else x.bfloat16()
    gamma, beta = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
    y, stats = ops.layernorm_fwd(x, gamma, beta)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (D,), gr, br)
    assert rel_err(y, yr) < 1e-2
    dy = torch.randn(M, D, device=DEV).bfloat16()
    skip = torch.randn(M, D, device=DEV).bfloat16()
    yr.backward(dy.float())
    dg, db, dc = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    dx = ops.layernorm_bwd(dy, x, stats, gamma, dgamma=dg, dbeta=db, skip=skip, dx_colsum=dc)
    assert rel_err(dx, xr.grad + skip.float()) < 1e-2
    assert rel_err(dg, gr.grad) < 1e-3 and rel_err(db, br.grad) < 1e-3
    assert rel_err(dc, (xr.grad + skip.float()).sum(0)) < 1e-3


@pytest.mark.parametrize("M", [1, 3, 9, 2049, 4737])
def test_layernorm_bwd_ragged_rows(ops, M):
    """Row counts that leave the second row of a warp's two-row iteration (and whole ring stages) unloaded: the
    parameter-gradient sums must not pick up anything from stale / never-written shared memory (NaN poison first)."""
    torch.manual_seed(M)
    D = 256
    poison = torch.full((148 * 2 * 8 * 4 * 3, D), float("nan"), device=DEV).bfloat16()    # touch lots of memory with NaN
    del poison
    x = (torch.randn(M, D, device=DEV) * 2 + 0.5).bfloat16()
    gamma, beta = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
    y, stats = ops.layernorm_fwd(x, gamma, beta)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    F.layer_norm(xr, (D,), gr, br).backward((dy := torch.randn(M, D, device=DEV).bfloat16()).float())
    dg, db, dc = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    dx = ops.layernorm_bwd(dy, x, stats, gamma, dgamma=dg, dbeta=db, dx_colsum=dc)
    assert torch.isfinite(dx.float()).all() and torch.isfinite(dg).all() and torch.isfinite(db).all() and torch.isfinite(dc).all()
    assert rel_err(dx, xr.grad) < 1e-2 and rel_err(dg, gr.grad) < 2e-3 and rel_err(db, br.grad) < 2e-3
    assert rel_err(dc, xr.grad.sum(0)) < 2e-3 + 1e-2 * (M < 4)


def test_layernorm_row_remap_and_adds(ops):
    torch.manual_seed(6)
    M, D = 64, 256
    x = torch.randn(M, D, device=DEV).bfloat16()
    gamma, beta = torch.randn(D, device=DEV), torch.randn(D, device=DEV)
    dst = torch.full((M,), -1, dtype=torch.int32, device=DEV)
    keep = torch.randperm(M)[:40]
    dst[keep] = torch.arange(40, dtype=torch.int32, device=DEV)
    a0, a1 = torch.randn(3, D, device=DEV), torch.randn(M, D, device=DEV)
    r0 = torch.randint(0, 3, (M,), dtype=torch.int32, device=DEV)
    r1 = torch.randperm(M).to(torch.int32).to(DEV)
    y, _ = ops.layernorm_fwd(x, gamma, beta, out_rows=40, dst_row=dst, add0=a0, add0_row=r0, add1=a1, add1_row=r1)
    ref = F.layer_norm(x.float(), (D,), gamma, beta) + a0[r0.long()] + a1[r1.long()]
    assert rel_err(y, ref[keep.to(DEV)]) < 1e-2
    # backward gather: rows without a source get zero dy
    dy = torch.randn(40, D, device=DEV).bfloat16()
    stats = torch.stack([x.float().mean(1), (x.float().var(1, unbiased=False) + 1e-5).rsqrt()], 1).contiguous()
    dx = ops.layernorm_bwd(dy, x, stats, gamma, src_row=dst)
    xr = x.float().requires_grad_(True)
    full = torch.zeros(M, D, device=DEV)
    full[keep.to(DEV)] = dy.float()
    F.layer_norm(xr, (D,), gamma, beta).backward(full)
    assert rel_err(dx, xr.grad) < 1e-2


# ------------------------------------------------------------------------------------ attention
def _attn_ref(qkv, B, n, H, scale):
    q, k, v = qkv.float().reshape(B, n, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * scale
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * n, H * 64), torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,n,H", [(3, 10, 4), (5, 4, 4), (2, 50, 4), (3, 64, 2), (2, 192, 4), (1, 256, 1), (2, 130, 3),
                                   # persistent kernels: several (sample, head) items per CTA
                                   (96, 192, 4), (120, 10, 4), (45, 256, 4), (50, 100, 8)])
def test_attention_fwd_bwd(ops, B, n, H):
    torch.manual_seed(7)
    scale = 64 ** -0.5
    qkv = torch.randn(B * n, 3 * H * 64, device=DEV).bfloat16()
    out, lse = ops.attention_fwd(qkv, B, n, H, 64, scale)
    qr = qkv.float().requires_grad_(True)
    oref, lref = _attn_ref(qr, B, n, H, scale)
    assert rel_err(out, oref) < 2e-2
    assert (lse - lref).abs().max().item() < 2e-2
    dout = torch.randn(B * n, H * 64, device=DEV).bfloat16()
    oref.backward(dout.float())
    dqkv = ops.attention_bwd(qkv, out, dout, lse, B, n, H, 64, scale)
    inner = H * 64
    for name, sl in (("dq", slice(0, inner)), ("dk", slice(inner, 2 * inner)), ("dv", slice(2 * inner, 3 * inner))):
        assert cos(dqkv[:, sl], qr.grad[:, sl]) > 0.999, name
        assert rel_err(dqkv[:, sl], qr.grad[:, sl]) < 5e-2, name
    # same with delta = rowsum(dO * O) supplied (the product path: fused into the GEMM that makes dO)
    delta = (dout.float() * out.float()).reshape(B * n, H, 64).sum(-1).contiguous()
    dqkv2 = ops.attention_bwd(qkv, out, dout, lse, B, n, H, 64, scale, delta=delta)
    assert rel_err(dqkv2, dqkv) < 1e-2


# ------------------------------------------------------------------- decoder assembly / sums
def test_decoder_assemble_fwd_bwd(ops):
    torch.manual_seed(8)
    B, n, D, nv = 5, 192, 256, 10
    noise = torch.rand(B, n)
    segs = [(0, 64, 60), (64, 64, 61), (128, 64, 61)]
    masked, unmasked, slots = ops.mask_indices(noise.to(DEV), segs)
    d = torch.randn(B * nv, D, device=DEV).bfloat16()
    mask_token = torch.randn(D, device=DEV)
    mod = torch.randn(3, D, device=DEV)
    pos = torch.randn(n, D, device=DEV)
    cls = (torch.arange(n) // 64).to(torch.int32).to(DEV)
    z = ops.decoder_assemble_fwd(d, nv, mask_token, slots, B, n, add0=mod, tok_class=cls, add1=pos)
    br = torch.arange(B, device=DEV)[:, None]
    ref = torch.zeros(B, n, D, device=DEV)
    ref[br, unmasked] = d.float().reshape(B, nv, D)
    ref[br, masked] = mask_token
    ref = ref + mod[cls.long()] + pos
    assert rel_err(z, ref.reshape(B * n, D)) < 1e-2
    dz = torch.randn(B * n, D, device=DEV).bfloat16()
    dmask, dmod, dpos = torch.zeros(D, device=DEV), torch.zeros(3, D, device=DEV), torch.zeros(n, D, device=DEV)
    dd = ops.decoder_assemble_bwd(dz, slots, B, n, nv, dmask_token=dmask, dadd0=dmod, tok_class=cls, dadd1=dpos)
    dzf = dz.float().reshape(B, n, D)
    assert torch.equal(dd.reshape(B, nv, D), dz.reshape(B, n, D)[br, unmasked])
    assert rel_err(dmask, dzf[br, masked].sum((0, 1))) < 1e-4
    assert rel_err(dmod, torch.stack([dzf[:, 64 * i:64 * (i + 1)].sum((0, 1)) for i in range(3)])) < 1e-4
    assert rel_err(dpos, dzf.sum(0)) < 1e-4


def test_rowclass_colsum_lnparam(ops):
    torch.manual_seed(9)
    B, nv, D = 33, 10, 256
    dx = torch.randn(B * nv, D, device=DEV).bfloat16()
    cls = torch.tensor([0, 0, 0, 0, 1, 1, 1, 2, 2, 2], dtype=torch.int32, device=DEV)
    pos = torch.randint(0, 192, (B * nv,), dtype=torch.int32, device=DEV)
    dcls, dpos = torch.zeros(3, D, device=DEV), torch.zeros(192, D, device=DEV)
    ops.rowclass_sum(dx, B, nv, slot_class=cls, dclass=dcls, row_pos=pos, dpos=dpos)
    f = dx.float().reshape(B, nv, D)
    assert rel_err(dcls, torch.stack([f[:, :4].sum((0, 1)), f[:, 4:7].sum((0, 1)), f[:, 7:].sum((0, 1))])) < 1e-4
    ref = torch.zeros(192, D, device=DEV).index_add_(0, pos.long(), dx.float())
    assert rel_err(dpos, ref) < 1e-4
    x = torch.randn(5000, 776, device=DEV).bfloat16()
    out = torch.ones(768, device=DEV)
    ops.colsum(x[:, :768], out)
    assert rel_err(out, 1 + x[:, :768].float().sum(0)) < 1e-4
    da, xh = torch.randn(1030, 192, device=DEV).bfloat16(), torch.randn(1030, 192, device=DEV).bfloat16()
    dg, db = torch.zeros(192, device=DEV), torch.zeros(192, device=DEV)
    ops.ln_param_grad(da, xh, dg, db)
    assert rel_err(dg, (da.float() * xh.float()).sum(0)) < 1e-4 and rel_err(db, da.float().sum(0)) < 1e-4


@pytest.mark.parametrize("C,H,W,k,st,pd", [(12, 64, 64, 4, 2, 1), (32, 16, 16, 4, 2, 1), (64, 8, 8, 3, 1, 1), (128, 8, 8, 1, 1, 0)])
def test_conv_stem_pieces(ops, C, H, W, k, st, pd):
    """im2col + GEMM(+bias+ReLU) == Conv2d + ReLU; col2im_relu + GEMMs == its backward (pretrain_models.py:37-56)."""
    torch.manual_seed(13)
    B, cout = 3, 64
    x = torch.randn(B, C, H, W, device=DEV)
    conv = torch.nn.Conv2d(C, cout, k, stride=st, padding=pd).to(DEV)
    wb = conv.weight.detach().bfloat16()
    xq = x.bfloat16().float()                                   # the kernels see bf16 activations
    ho, wo = ops.conv_out_size(H, k, st, pd), ops.conv_out_size(W, k, st, pd)
    col = ops.im2col(x, B, C, H, W, k, st, pd, False)
    ref_col = F.unfold(xq, k, padding=pd, stride=st).transpose(1, 2).reshape(B * ho * wo, C * k * k)
    assert torch.equal(col.float(), ref_col.bfloat16().float())
    x_nhwc = xq.permute(0, 2, 3, 1).reshape(B * H * W, C).bfloat16().contiguous()
    assert torch.equal(ops.im2col(x_nhwc, B, C, H, W, k, st, pd, True), col)
    y = ops.gemm(col, wb.reshape(cout, -1).contiguous(), bias=conv.bias.detach(), act=ops.RELU)
    xr = xq.clone().requires_grad_(True)
    yr = F.relu(F.conv2d(xr, wb.float(), conv.bias, stride=st, padding=pd))
    assert rel_err(y, yr.permute(0, 2, 3, 1).reshape(B * ho * wo, cout)) < 1e-2
    # backward of the convolution input (dgrad GEMM + gather) fused with a ReLU mask of the layer below
    dy = torch.randn(B * ho * wo, cout, device=DEV).bfloat16()
    mask_src = torch.randn(B * H * W, C, device=DEV).bfloat16()
    dcol = ops.gemm(dy, wb.reshape(cout, -1).T.contiguous())
    dx = ops.col2im_relu(dcol, B, C, H, W, k, st, pd, relu_out=mask_src)
    g = torch.autograd.grad(F.conv2d(xr, wb.float(), None, stride=st, padding=pd),
                            xr, dy.float().reshape(B, ho, wo, cout).permute(0, 3, 1, 2))[0]
    ref_dx = g.permute(0, 2, 3, 1).reshape(B * H * W, C) * (mask_src.float() > 0)
    assert rel_err(dx, ref_dx) < 2e-2
    assert rel_err(ops.col2im_relu(dcol, B, C, H, W, k, st, pd), g.permute(0, 2, 3, 1).reshape(B * H * W, C)) < 2e-2


def test_token_finish_fwd_bwd(ops):
    torch.manual_seed(14)
    B, n_per, nsrc, D, n_total, tok_base, nv = 5, 16, 2, 128, 48, 16, 6
    x = torch.randn(nsrc * B * n_per, D, device=DEV).bfloat16()
    tok = torch.stack([torch.randperm(nsrc * n_per)[:nv] + tok_base for _ in range(B)]).to(DEV).to(torch.int32)
    idx = torch.zeros(B, 9, dtype=torch.int32, device=DEV)
    idx[:, 3:] = tok
    add0, add1 = torch.randn(3, D, device=DEV), torch.randn(n_total, D, device=DEV)
    cls = torch.tensor([0] * 16 + [1] * 16 + [2] * 16, dtype=torch.int32, device=DEV)
    dst = (torch.arange(B, device=DEV)[:, None] * 10 + 2 + torch.arange(nv, device=DEV)[None]).reshape(-1).to(torch.int32)
    out = torch.zeros(B * 10, D, device=DEV, dtype=torch.bfloat16)
    ops.token_finish(x, B, n_per, nv, tok_base, out, tok_idx=idx, col0=3, add0=add0, tok_class=cls, add1=add1, dst_row=dst)
    xs = x.float().reshape(nsrc, B, n_per, D)
    for b in range(B):
        for j in range(nv):
            t = int(tok[b, j]); tl = t - tok_base
            ref = xs[tl // n_per, b, tl % n_per] + add0[int(cls[t])] + add1[t]
            assert rel_err(out[b * 10 + 2 + j], ref) < 1e-2
    # backward: gather with a slot map (-1 = masked token -> zero row)
    slots = torch.full((B, n_total), -1, dtype=torch.int32, device=DEV)
    for b in range(B):
        for j in range(nv):
            slots[b, int(tok[b, j])] = 2 + j
    dx0 = torch.randn(B * 10, D, device=DEV).bfloat16()
    dtok = ops.token_finish_bwd(dx0, B, 10, n_total, tok_base, nsrc * n_per, n_per, slot_of_token=slots)
    ref = torch.zeros(nsrc, B, n_per, D, device=DEV)
    for b in range(B):
        for j in range(nv):
            tl = int(tok[b, j]) - tok_base
            ref[tl // n_per, b, tl % n_per] = dx0[b * 10 + 2 + j].float()
    assert torch.equal(dtok.float(), ref.reshape(-1, D))


def test_token_mean(ops):
    torch.manual_seed(12)
    B, n, D = 37, 192, 256
    x = torch.randn(B * n, D, device=DEV).bfloat16()
    out = ops.token_mean_fwd(x, B, n)
    assert rel_err(out, x.float().reshape(B, n, D).mean(1)) < 1e-5
    g = torch.randn(B, D, device=DEV)
    dx = ops.token_mean_bwd(g, B, n)
    assert rel_err(dx, (g / n)[:, None, :].expand(B, n, D).reshape(B * n, D)) < 1e-2


def test_mse_loss_and_grad(ops):
    torch.manual_seed(10)
    B, nm = 7, 60
    img = torch.rand(B, 12, 64, 64, device=DEV)
    ps = ops.make_patch_source([img], 8, 8, token_base=0)
    idx = torch.stack([torch.randperm(64)[:nm] for _ in range(B)]).to(DEV)
    pred = torch.randn(B * nm, 768, device=DEV)
    acc = torch.zeros(1, device=DEV)
    w = 1.0 / pred.numel()
    dpred = ops.mse_loss(ps, B, nm, pred, w, acc, tok_idx=idx)
    tgt = _patchify(img, 8, 8)[torch.arange(B, device=DEV)[:, None], idx].reshape(B * nm, 768)
    pr = pred.clone().requires_grad_(True)
    l = F.mse_loss(pr, tgt)
    l.backward()
    assert abs(acc.item() - l.item()) < 1e-5 * abs(l.item())
    assert rel_err(dpred, pr.grad) < 1e-2
    # fused bias gradient (column sums of dpred) + tactile geometry (two sensors, 4x4 patches, ragged grid)
    cs = torch.ones(768, device=DEV)
    acc2 = torch.zeros(1, device=DEV)
    dpred2 = ops.mse_loss(ps, B, nm, pred, w, acc2, tok_idx=idx, dpred_colsum=cs)
    assert torch.equal(dpred2, dpred) and acc2.item() == acc.item()       # deterministic
    assert rel_err(cs, 1 + pr.grad.sum(0)) < 1e-4
    t1, t2 = torch.rand(B, 12, 32, 32, device=DEV), torch.rand(B, 12, 32, 32, device=DEV)
    pst = ops.make_patch_source([t1, t2], 4, 4, token_base=64)
    nmt = 122
    idt = torch.stack([torch.randperm(128)[:nmt] + 64 for _ in range(B)]).to(DEV)
    predt = torch.randn(B * nmt, 192, device=DEV)
    acct, cst = torch.zeros(1, device=DEV), torch.zeros(192, device=DEV)
    wt = 10.0 / predt.numel()
    dpt = ops.mse_loss(pst, B, nmt, predt, wt, acct, tok_idx=idt, dpred_colsum=cst)
    tg = torch.cat([_patchify(t1, 4, 4), _patchify(t2, 4, 4)], 1)[torch.arange(B, device=DEV)[:, None], idt - 64].reshape(B * nmt, 192)
    prt = predt.clone().requires_grad_(True)
    lt = 10 * F.mse_loss(prt, tg)
    lt.backward()
    assert abs(acct.item() - lt.item()) < 1e-5 * abs(lt.item())
    assert rel_err(dpt, prt.grad) < 1e-2 and rel_err(cst, prt.grad.sum(0)) < 1e-4


# ------------------------------------------------------------------------------------ optimizer
def test_clip_adamw_matches_torch(ops):
    torch.manual_seed(11)
    n = 100003
    p0 = torch.randn(n, device=DEV)
    p = p0.clone()
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p_ref], lr=1e-3)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    state = torch.zeros(ops.OPT_STATE_DOUBLES, dtype=torch.float64, device=DEV)
    for it in range(3):
        g = torch.randn(n, device=DEV) * (3.0 if it < 2 else 1e-4)  # clip active, then inactive
        p_ref.grad = g.clone()
        total_ref = torch.nn.utils.clip_grad_norm_([p_ref], 0.5)
        opt.step()
        gg = g.clone()
        state[1] = 0
        ops.grad_sumsq(gg, state)
        ops.optimizer_step_begin(state)
        ops.clip_adamw(p, gg, m, v, state, lr=1e-3)
        assert abs(state[2].item() - total_ref.item()) < 1e-5 * total_ref.item()
        assert rel_err(gg, p_ref.grad) < 1e-5
        assert rel_err(p, p_ref.data) < 1e-5
    assert state[0].item() == 3.0
    # hyper-parameters read from a device buffer at run time (what a captured graph does) give the same step as scalars
    hyper = torch.tensor([2e-3, 0.8, 0.95, 1e-7, 0.1, 0.25, 0, 0], device=DEV)
    opt2 = torch.optim.AdamW([p_ref], lr=2e-3, betas=(0.8, 0.95), eps=1e-7, weight_decay=0.1)
    opt2.load_state_dict({"state": opt.state_dict()["state"], "param_groups": opt2.state_dict()["param_groups"]})
    g = torch.randn(n, device=DEV)
    p_ref.grad = g.clone()
    torch.nn.utils.clip_grad_norm_([p_ref], 0.25)
    opt2.step()
    state[1] = 0
    ops.grad_sumsq(g, state)
    ops.optimizer_step_begin(state, (0.0, 0.0), hyper=hyper)          # the scalar arguments are ignored
    ops.clip_adamw(p, g, m, v, state, lr=123.0, max_norm=99.0, hyper=hyper)
    assert rel_err(p, p_ref.data) < 1e-5 and state[0].item() == 4.0


def test_cast_and_transpose(ops):
    from m3l_b200 import _lib
    torch.manual_seed(12)
    src = torch.randn(3000, device=DEV)
    dst = torch.empty(3000, dtype=torch.bfloat16, device=DEV)
    ops.cast_bf16(src, dst)
    assert torch.equal(dst, src.bfloat16())
    base = torch.randn(768 * 256 + 100 * 52, device=DEV)
    out = torch.zeros(base.numel(), dtype=torch.bfloat16, device=DEV)
    descs = (_lib.MatrixDesc * 2)()
    descs[0].src_offset, descs[0].dst_offset, descs[0].rows, descs[0].cols = 0, 0, 768, 256
    descs[1].src_offset, descs[1].dst_offset, descs[1].rows, descs[1].cols = 768 * 256, 768 * 256, 100, 52
    dd = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(DEV)
    ops.transpose_cast_bf16(base, out, dd, 2)
    assert torch.equal(out[:768 * 256].reshape(256, 768), base[:768 * 256].reshape(768, 256).T.bfloat16())
    assert torch.equal(out[768 * 256:].reshape(52, 100), base[768 * 256:].reshape(100, 52).T.bfloat16())


def test_errors_are_loud(ops):
    from m3l_b200._lib import M3LError
    a = torch.randn(64, 64, device=DEV).bfloat16()
    b = torch.randn(60, 64, device=DEV).bfloat16()  # N = 60 is not a multiple of 8
    with pytest.raises(M3LError):
        ops.gemm(a, b)
    qkv = torch.randn(2 * 300, 3 * 64, device=DEV).bfloat16()  # n = 300 > 256
    with pytest.raises(M3LError):
        ops.attention_fwd(qkv, 2, 300, 1, 64, 0.125)


@pytest.mark.parametrize("M,K,with_skip", [(49152, 1024, True), (2560, 512, True), (1000, 768, False), (128, 64, True)])
def test_gemm_layernorm_backward_epilogue(ops, M, K, with_skip):
    """dgrad GEMM with the fused LayerNorm-backward epilogue (EPI_LN_BWD) against the two separate kernels it replaces
    (gemm -> m3l_layernorm_bwd) and against torch autograd of LayerNorm in fp32."""
    torch.manual_seed(3)
    D = 256
    a = (torch.randn(M, K, device=DEV) * 0.5).bfloat16()
    w = (torch.randn(D, K, device=DEV) * (K ** -0.5)).bfloat16()
    x = torch.randn(M, D, device=DEV).bfloat16()
    gamma = torch.rand(D, device=DEV) + 0.5
    skip = torch.randn(M, D, device=DEV).bfloat16() if with_skip else None
    _, stats = ops.layernorm_fwd(x, gamma, torch.zeros(D, device=DEV), want_stats=True)
    # separate kernels
    dy = ops.gemm(a, w)
    dg0, db0, dc0 = (torch.zeros(D, device=DEV) for _ in range(3))
    dx0 = ops.layernorm_bwd(dy, x, stats, gamma, dgamma=dg0, dbeta=db0, skip=skip, dx_colsum=dc0)
    # fused
    dg1, db1, dc1 = (torch.zeros(D, device=DEV) for _ in range(3))
    dx1 = ops.gemm(a, w, ln_bwd=dict(x=x, stats=stats, gamma=gamma, skip=skip, dgamma=dg1, dbeta=db1, dx_colsum=dc1))
    assert dx1.shape == (M, D) and dx1.dtype == torch.bfloat16
    assert cos(dx1, dx0) > 0.9999 and rel_err(dx1, dx0) < 2e-2
    for name, u, v in (("dgamma", dg1, dg0), ("dbeta", db1, db0), ("dx_colsum", dc1, dc0)):
        assert cos(u, v) > 0.9999 and rel_err(u, v) < 2e-2, name
    # fp32 autograd statement
    xr = x.float().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    br = torch.zeros(D, device=DEV, requires_grad=True)
    y = F.layer_norm(xr, (D,), gr, br)
    dyr = a.float() @ w.float().T
    y.backward(dyr)
    ref = xr.grad + (skip.float() if with_skip else 0)
    assert cos(dx1, ref) > 0.9995
    assert cos(dg1, gr.grad) > 0.9995 and cos(db1, br.grad) > 0.9995
    assert cos(dc1, ref.sum(0)) > 0.999
