"""Autograd side of the explainer: what ``temp_exp_main.py``'s training loop needs from the
drop-in -- gradients of the motif scores, of ``retrieve_explanation`` (Beta ``rsample``) and of ``kl_loss`` with respect
to the explainer's parameters (reference models/explainer.py:174-201, 354-453; loop temp_exp_main.py:605-632).

Division of labour
  * forward values without dropout (eval mode, ``dropout_p == 0``, or no parameter requires grad) always come from the fused
    sm_100a scorer; ``FusedScore`` attaches a backward that recomputes the layers from the saved walk tensors;
  * the layer-by-layer differentiable evaluation below (``scores_layerwise``) is the recompute of that backward and the
    training-mode forward (dropout masks must be shared by forward and backward); every nn.Linear in it goes through ``TcLinear``:
    forward, dgrad and wgrad products on the library's tcgen05 3xTF32 GEMM (``tm_gemm_tf32x3``), the elementwise glue is torch;
  * hand-written kernels with their own backward: the Beta sampler (``tm_beta_sample``: Philox + Marsaglia-Tsang gammas, pathwise
    gradient through the two gammas) and ``kl_loss`` (``tm_kl_loss`` / ``tm_kl_loss_backward``).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from ._lib import check, lib, ptr


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


# ---------------------------------------------------------------------------------------------------- Linear on the tcgen05 GEMM
def gemm(a, b, bias=None, out=None, accumulate=False):
    """out[M, N] (+)= a[M, K] @ b[N, K]^T (+ bias[N]) on tm_gemm_tf32x3 (fp32 in / out, 3xTF32 on the tensor cores)."""
    a = a.contiguous(); b = b.contiguous()
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K and a.dtype == torch.float32 and b.dtype == torch.float32
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    bs = bias.contiguous() if bias is not None else None
    with torch.cuda.device(a.device):
        check(lib().tm_gemm_tf32x3(M, N, K, ptr(a), K, ptr(b), K, ptr(out), out.stride(0), ptr(bs), int(bool(accumulate)), _stream(a.device)), "tm_gemm_tf32x3")
    return out


class TcLinear(torch.autograd.Function):
    """y = x W^T + b with forward, dgrad (dx = dy W) and wgrad (dW = dy^T x) on the tensor cores."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x2 = x.reshape(-1, x.shape[-1]).to(torch.float32)
        ctx.save_for_backward(x2, weight)
        ctx.shape = x.shape
        ctx.has_bias = bias is not None
        y = gemm(x2, weight.detach(), bias.detach() if bias is not None else None)
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, weight = ctx.saved_tensors
        g2 = gy.reshape(-1, gy.shape[-1]).to(torch.float32).contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = gemm(g2, weight.detach().t().contiguous()).view(ctx.shape)            # dX = dY W: B operand = W^T [K, N]
        if ctx.needs_input_grad[1]:
            gw = gemm(g2.t().contiguous(), x2.t().contiguous())                        # dW = dY^T X: reduction over the rows
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = g2.sum(0)
        return gx, gw, gb


def _linear(mod, x):
    return TcLinear.apply(x, mod.weight, mod.bias)


def _seq(seq, x):
    """nn.Sequential of Linear / ReLU / Dropout with the Linears on TcLinear."""
    for layer in seq:
        x = _linear(layer, x) if isinstance(layer, torch.nn.Linear) else layer(x)
    return x


# ---------------------------------------------------------------------------------------------------- layer-by-layer scorer
def _time_encode(m, dt):
    """TimeEncode (explainer.py:51-59): cos(dt * basis_freq + phase), dt [...]. -> [..., D]"""
    te = m.time_encoder
    return torch.cos(dt.unsqueeze(-1) * te.basis_freq + te.phase)


def attention_layerwise(m, nodes, eidx, t, cut, eid):
    """Walk embeddings [B, W, H] = attention(event_conv(...)) (explainer.py:175-193), differentiable in the module's parameters.
    nodes i64 [B,W,6] (src3,tgt3,src2,tgt2,src1,tgt1), eidx i64 [B,W,3], t f32 [B,W,3], cut f32 [B], eid f32 [B,W,3,3]."""
    ec, att = m.event_conv, m.attention
    x = torch.cat([m.edge_raw_embed(eidx), eid, _time_encode(m, t[..., 2:3] - t)], dim=-1)       # [B,W,3,ev] (:176-179, :326)
    ev = _linear(ec.lin_event, x)
    a, b = m.node_raw_embed(nodes[..., 0::2]), m.node_raw_embed(nodes[..., 1::2])                # the events' two endpoints (:348-351)
    h = torch.cat([_seq(ec.MLP, a + torch.relu(b + ev)), _seq(ec.MLP, b + torch.relu(a + ev))], dim=-1)      # both orientations (:182-186)
    q, k = h[:, :, 2], h[:, :, :2]                                                               # the event next to the root queries the other two
    wp, wq = _linear(att.W1, q), _linear(att.W2, k)                                              # [B,W,2H], [B,W,2,2H]
    s = (wq * wp.unsqueeze(2)).sum(-1)                                                           # [B,W,2]
    if m.use_temporal_guidance:
        td = (cut.view(-1, 1, 1) - t[..., :2]).abs()
        s = s * (0.7 + 0.3 * torch.exp(-td / (td.std() + 1e-6)))                                 # batch-global std (:826-836)
    alpha = torch.softmax(s, dim=-1)
    if m.use_temporal_guidance:
        alpha = att.dropout(alpha)
    return _seq(att.MLP, q + (alpha.unsqueeze(-1) * wq).sum(2))                                  # (:841-843)


def scores_layerwise(m, nodes, eidx, t, cat, cut, eid):
    """Motif scores [B, W, 1] (explainer.py:174-201)."""
    y = attention_layerwise(m, nodes, eidx, t, cut, eid)
    if m.if_cat:
        y = torch.cat([y, F.one_hot(cat.long(), 12).to(y.dtype)], dim=-1)
    return torch.sigmoid(_seq(m.MLP, y))


class FusedScore(torch.autograd.Function):
    """scores = fused kernel (tm_encode_score); backward = recompute of the layers from the saved walk tensors."""

    @staticmethod
    def forward(ctx, m, nodes, eidx, t, cat, cut, eid, *params):
        ctx.m = m
        ctx.inputs = (nodes, eidx, t, cat, cut, eid)
        with torch.no_grad():
            s = m.score_device(nodes, eidx, t, cat, cut, eid)
        return s.view(nodes.shape[0], nodes.shape[1], 1)

    @staticmethod
    def backward(ctx, grad):
        m = ctx.m
        nodes, eidx, t, cat, cut, eid = ctx.inputs
        params = [p for p in m.scorer_parameters() if p.requires_grad]
        with torch.enable_grad():
            out = scores_layerwise(m, nodes.long(), eidx.long(), t, cat, cut, eid)
            grads = torch.autograd.grad(out, params, grad, allow_unused=True)
        it = iter(grads)
        return (None,) * 7 + tuple(next(it) if p.requires_grad else None for p in m.scorer_parameters())


def score_autograd(m, nodes, eidx, t, cat, cut, eid):
    """Scores with an autograd graph.  Active dropout (training mode, p > 0): the layer-by-layer evaluation end to end; otherwise
    the fused kernel's values with the recompute backward."""
    if m._dropout_active():
        return scores_layerwise(m, nodes.long(), eidx.long(), t, cat, cut, eid)
    return FusedScore.apply(m, nodes, eidx, t, cat, cut, eid, *m.scorer_parameters())


# ---------------------------------------------------------------------------------------------------- Beta rsample
class BetaRSample(torch.autograd.Function):
    """x ~ Beta(max(10 p, 1), max(10 (1 - p), 1)) (explainer.py:421-427) on the device; dx/dp through the two gammas:
    x = g1 / (g1 + g2), dg/da = torch._standard_gamma_grad (the implicit reparameterisation of the gamma)."""

    @staticmethod
    def forward(ctx, prob, node, seed, offset):
        p = prob.detach().to(torch.float32).contiguous()
        out = torch.empty_like(p); g1 = torch.empty_like(p); g2 = torch.empty_like(p)
        nd = node.to(torch.int32).contiguous() if node is not None else None
        with torch.cuda.device(p.device):
            check(lib().tm_beta_sample(p.numel(), ptr(p), ptr(nd), int(seed), int(offset), ptr(out), ptr(g1), ptr(g2), _stream(p.device)), "tm_beta_sample")
        ctx.save_for_backward(p, g1, g2)
        ctx.mask = None if nd is None else (nd == 0)
        return out

    @staticmethod
    def backward(ctx, grad):
        p, g1, g2 = ctx.saved_tensors
        a10, b10 = p * 10, (1 - p) * 10
        alpha, beta = a10.clamp(min=1.0), b10.clamp(min=1.0)
        tot2 = (g1 + g2) ** 2
        dx_da = g2 / tot2 * torch._standard_gamma_grad(alpha, g1)
        dx_db = -g1 / tot2 * torch._standard_gamma_grad(beta, g2)
        gp = grad * (dx_da * 10 * (a10 >= 1) - dx_db * 10 * (b10 >= 1))
        if ctx.mask is not None:
            gp = gp.masked_fill(ctx.mask, 0)
        return gp, None, None, None


def next_seed():
    """Philox key of one sampling call, drawn from torch's CPU generator: torch.manual_seed makes the draws reproducible."""
    return int(torch.randint(0, 2 ** 62, (1,)).item())


# ---------------------------------------------------------------------------------------------------- motif -> edge aggregation
def edge_importance_autograd(m, scores, eidx_w, t_w, h_nodes, h_eidx, training):
    """retrieve_edge_imp_node (explainer.py:354-406) with an autograd graph: dependency gate, per-root max over the walks that carry
    an edge id, gather to the hop slots, Beta rsample / mean, padding mask.  scores [B,W,1] (graph attached), eidx_w i64 [B,W,3],
    t_w f32 [B,W,3], h_nodes / h_eidx: lists of i64 [B,K_l].  The reference scatters into a dense [B, max edge id + 1] matrix; here the
    (root, edge id) pairs are compacted with a sort so that memory does not grow with the graph."""
    B, W = eidx_w.shape[0], eidx_w.shape[1]
    e = eidx_w.reshape(B, 3 * W)
    imp = scores.reshape(B, W, 1).expand(B, W, 3).reshape(B, 3 * W)                               # graphlet_imp.repeat(1,1,3) (:363)
    if m.use_dependency_aware_sampling:
        feat = torch.cat([m.edge_raw_embed(e), _time_encode(m, t_w.reshape(B, 3 * W))], dim=-1)  # raw timestamps (:371-375)
        gate = torch.sigmoid(_seq(m.edge_dependency_gcn, feat).squeeze(-1))
        imp = imp * (0.5 + 0.5 * gate)                                                           # (:383-386)
    stride = int(max(int(e.max()), max(int(x.max()) for x in h_eidx))) + 1
    row = torch.arange(B, device=e.device).view(B, 1) * stride
    uniq, inv = torch.unique((row + e).reshape(-1), return_inverse=True)
    best = torch.zeros(uniq.numel(), dtype=imp.dtype, device=imp.device).scatter_reduce(0, inv, imp.reshape(-1), "amax", include_self=True)
    outs = []
    for l, (nd, ee) in enumerate(zip(h_nodes, h_eidx)):
        key = (row + ee).reshape(-1)
        pos = torch.searchsorted(uniq, key).clamp(max=uniq.numel() - 1)
        p = torch.where(uniq[pos] == key, best[pos], torch.zeros((), dtype=imp.dtype, device=imp.device)).view(ee.shape)   # ids no walk carries: 0 (:389)
        if training:
            x = BetaRSample.apply(p, nd, next_seed(), l << 40)
        else:
            alpha, beta = (p * 10).clamp(min=1.0), ((1 - p) * 10).clamp(min=1.0)
            x = (alpha / (alpha + beta)).masked_fill(nd == 0, 0)
        outs.append(x)
    return outs


# ---------------------------------------------------------------------------------------------------- kl_loss
class KLLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prob, cat, null, target, empirical):
        B, W = cat.shape
        p = prob.detach().to(torch.float32).reshape(B, W).contiguous()
        work = torch.empty(B, dtype=torch.float64, device=p.device)
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            check(lib().tm_kl_loss(B, W, ptr(p), ptr(cat), ptr(null), int(null.numel()), float(target), int(empirical), ptr(work), ptr(loss), _stream(p.device)),
                  "tm_kl_loss")
        ctx.save_for_backward(p, cat, null)
        ctx.args = (float(target), int(empirical), prob.shape)
        return loss

    @staticmethod
    def backward(ctx, grad):
        p, cat, null = ctx.saved_tensors
        target, empirical, shape = ctx.args
        B, W = cat.shape
        g = grad.detach().to(torch.float32).contiguous()
        gp = torch.empty_like(p)
        with torch.cuda.device(p.device):
            check(lib().tm_kl_loss_backward(B, W, ptr(p), ptr(cat), ptr(null), int(null.numel()), target, empirical, ptr(g), ptr(gp), _stream(p.device)),
                  "tm_kl_loss_backward")
        return gp.view(shape), None, None, None, None
