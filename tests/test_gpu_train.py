"""GPU tests of the training side (-m gpu): one temp_exp_main.py:605-632-shaped step on the drop-in --
forward with an autograd graph, retrieve_edge_imp_node, kl_loss, backward -- against gradients recorded from the unmodified
reference (tests/golden/trainstep_*.npz, train() mode with dropout_p = 0); the device Beta sampler against the Beta law and its
analytic mean gradient; the reference's eval loop (training=True under eval()) on the fused kernels."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tm():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import tempme_b200
    return tempme_b200


class _Base:
    def __init__(self, nfeat, efeat):
        self.n_feat_th = torch.as_tensor(nfeat).cuda(); self.e_feat_th = torch.as_tensor(efeat).cuda()
        self.node_raw_features = torch.nn.Embedding.from_pretrained(self.n_feat_th, padding_idx=0, freeze=True)
        self.edge_raw_features = torch.nn.Embedding.from_pretrained(self.e_feat_th, padding_idx=0, freeze=True)


def model_from(tm, z, dropout_p=0.0):
    null = {k: float(v) for k, v in zip(range(1, 13), z["null_values"])}
    m = tm.TempME(_Base(z["node_feat"], z["edge_feat"]), "tgn", "unit", out_dim=40, hid_dim=int(z["hid_dim"]), prior=str(z["prior"]),
                  dropout_p=dropout_p, device="cuda", use_temporal_guidance=bool(z["use_temporal"]), null_model=null)
    sd = {k[2:]: torch.as_tensor(v) for k, v in z.items() if k.startswith("p:")}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected
    if dropout_p == 0.0:              # dropout_p does not reach the attention module (explainer.py:121): the goldens zero every Dropout
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
    return m.cuda()


def step_inputs(z):
    walks = (z["w_nodes"].astype(np.int64), z["w_eidx"].astype(np.int64), z["w_t"].astype(np.float64), z["w_cat"].astype(np.int64)[..., None], None)
    sub = ([z["h0_node"].astype(np.int64), z["h1_node"].astype(np.int64)], [z["h0_eidx"].astype(np.int64), z["h1_eidx"].astype(np.int64)], None)
    return walks, sub


@pytest.mark.parametrize("tag", ["d32", "d32_hid32_uniform", "d172"])
def test_train_step_matches_reference_gradients(tm, golden, tag):
    z = golden("trainstep_" + tag)
    m = model_from(tm, z).train()
    walks, sub = step_inputs(z)
    score = m(walks, z["cut_time"], z["edge_identity"].astype(np.float64))
    assert score.requires_grad and score.shape == z["score"].shape
    np.testing.assert_allclose(score.detach().cpu().numpy(), z["score"], rtol=1e-5, atol=0)        # values: the fused kernel
    imp0, imp1 = m.retrieve_edge_imp_node(sub, score, walks, training=False)
    np.testing.assert_allclose(imp0.detach().cpu().numpy(), z["imp0"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(imp1.detach().cpu().numpy(), z["imp1"], rtol=1e-5, atol=1e-7)
    kl = m.kl_loss(score, walks, target=0.3)
    np.testing.assert_allclose(float(kl), float(z["kl"]), rtol=2e-5)
    loss = (imp0 * torch.as_tensor(z["w0"]).cuda()).sum() + (imp1 * torch.as_tensor(z["w1"]).cuda()).sum() + 0.5 * kl
    np.testing.assert_allclose(float(loss), float(z["loss"]), rtol=1e-5)
    loss.backward()
    grads = {k: v.grad for k, v in m.named_parameters() if v.grad is not None}
    checked = 0
    for k in z:
        if not k.startswith("g:"):
            continue
        ref = z[k]
        assert k[2:] in grads, f"no gradient for {k[2:]}"
        g = grads[k[2:]].cpu().numpy()
        scale = max(float(np.abs(ref).max()), 1e-12)
        np.testing.assert_allclose(g, ref, rtol=1e-4, atol=2e-5 * scale, err_msg=k)
        checked += 1
    assert checked >= 20
    # an optimiser step changes what the fused kernel scores next (the packed weights follow the parameters' versions)
    opt = torch.optim.Adam(m.parameters(), lr=1e-2)
    opt.step()
    m.eval()
    with torch.no_grad():
        after = m(walks, z["cut_time"], z["edge_identity"].astype(np.float64))
    assert not after.requires_grad and float((after - score.detach()).abs().max()) > 1e-4


def test_training_mode_with_dropout_and_beta_rsample(tm, golden):
    """The reference's defaults: dropout_p = 0.1, retrieve_explanation(training=True) -> Beta rsample with a pathwise gradient."""
    z = golden("trainstep_d32")
    torch.manual_seed(0)
    m = model_from(tm, z, dropout_p=0.1).train()
    walks, sub = step_inputs(z)
    s = [m(walks, z["cut_time"], z["edge_identity"].astype(np.float64)) for _ in range(3)]
    assert not torch.equal(s[0], s[1])                                                              # dropout is live
    exp = m.retrieve_explanation(sub, s[0], walks, sub, s[1], walks, sub, s[2], walks, training=True)
    assert len(exp) == 2 and exp[0].shape == (36, z["h0_node"].shape[1]) and exp[1].shape == (36, z["h1_node"].shape[1])
    h0 = torch.as_tensor(np.concatenate([z["h0_node"]] * 3)).cuda()
    assert bool((exp[0][h0 == 0] == 0).all()) and bool(((exp[0] >= 0) & (exp[0] <= 1)).all())
    loss = exp[0].sum() + exp[1].mean() + sum(m.kl_loss(x, walks) for x in s)
    loss.backward()
    for name, p in m.named_parameters():
        if name.startswith(("event_conv", "attention.W1", "attention.W2", "attention.MLP", "MLP", "edge_dependency_gcn")):
            assert p.grad is not None and torch.isfinite(p.grad).all(), name
    assert float(m.event_conv.lin_event.weight.grad.abs().sum()) > 0 and float(m.edge_dependency_gcn[0].weight.grad.abs().sum()) > 0


def test_reference_eval_loop_samples_on_fused_kernels(tm, golden):
    """temp_exp_main.py's evaluation: explainer.eval(), gradients enabled, retrieve_* with training=args.if_bern (True)."""
    z = golden("trainstep_d32")
    m = model_from(tm, z, dropout_p=0.1).eval()
    walks, sub = step_inputs(z)
    score = m(walks, z["cut_time"], z["edge_identity"].astype(np.float64))
    assert not score.requires_grad
    np.testing.assert_allclose(score.cpu().numpy(), z["score"], rtol=1e-5, atol=0)
    torch.manual_seed(5)
    a0, a1 = m.retrieve_edge_imp_node(sub, score, walks)                                            # training defaults to True, as in the reference
    torch.manual_seed(5)
    b0, b1 = m.retrieve_edge_imp_node(sub, score, walks)
    assert torch.equal(a0, b0) and torch.equal(a1, b1)                                               # draws follow torch.manual_seed
    c0, _ = m.retrieve_edge_imp_node(sub, score, walks)
    assert not torch.equal(a0, c0)
    mean0, mean1 = m.retrieve_edge_imp_node(sub, score, walks, training=False)
    np.testing.assert_allclose(mean0.cpu().numpy(), z["imp0"], rtol=1e-5, atol=1e-7)
    live = torch.as_tensor(z["h1_node"]).cuda() != 0
    assert bool((a1[~live] == 0).all()) and bool(((a1[live] > 0) & (a1[live] < 1)).all())
    # many draws per slot: the sample mean approaches the Beta mean
    acc = torch.zeros_like(mean1)
    for _ in range(200):
        acc += m.retrieve_edge_imp_node(sub, score, walks)[1]
    err = (acc / 200 - mean1)[live].abs()
    assert float(err.mean()) < 0.02 and float(err.max()) < 0.12


def test_beta_sampler_law_and_gradient(tm):
    from scipy import stats
    from tempme_b200.training import BetaRSample
    n = 200_000
    for p0 in (0.03, 0.3, 0.5, 0.85, 0.999):
        p = torch.full((n,), p0, device="cuda", requires_grad=True)
        x = BetaRSample.apply(p, None, 1234, 0)
        a, b = max(10 * p0, 1.0), max(10 * (1 - p0), 1.0)
        ks = stats.kstest(x.detach().cpu().numpy().astype(np.float64), stats.beta(a, b).cdf)
        assert ks.pvalue > 1e-3, (p0, ks)
        # pathwise gradient: E[dx/dp] = d/dp E[x] = d/dp [alpha / (alpha + beta)]
        x.sum().backward()
        g = float(p.grad.mean())
        da, db = (10.0 if 10 * p0 >= 1 else 0.0), (-10.0 if 10 * (1 - p0) >= 1 else 0.0)
        want = (da * (a + b) - a * (da + db)) / (a + b) ** 2
        se = float(p.grad.std()) / np.sqrt(n)
        assert abs(g - want) < 6 * se + 1e-4, (p0, g, want, se)
    # different keys / offsets give different draws; the same key reproduces
    p = torch.rand(1000, device="cuda")
    assert torch.equal(BetaRSample.apply(p, None, 7, 0), BetaRSample.apply(p, None, 7, 0))
    assert not torch.equal(BetaRSample.apply(p, None, 7, 0), BetaRSample.apply(p, None, 8, 0))
    assert not torch.equal(BetaRSample.apply(p, None, 7, 0), BetaRSample.apply(p, None, 7, 1 << 40))


@pytest.mark.parametrize("prior", ["empirical", "uniform"])
def test_kl_loss_backward_matches_autograd_of_the_formula(tm, golden, prior):
    z = golden("kl_loss")
    null = torch.as_tensor(z["null_values"], dtype=torch.float32).cuda()

    class M:      # the two attributes kl_loss reads
        pass
    from tempme_b200.training import KLLoss
    for name in ("us", "few", "one"):
        cat = torch.as_tensor(z[f"{name}_cat"].astype(np.uint8)).cuda()
        B, W = cat.shape
        for target in (0.3, 0.05):
            p = torch.as_tensor(z[f"{name}_prob"]).cuda().reshape(B, W).clone().requires_grad_(True)
            loss = KLLoss.apply(p, cat, null, target, int(prior == "empirical"))
            np.testing.assert_allclose(float(loss), float(z[f"{name}_{prior}_{target}"]), rtol=2e-5, atol=1e-7)
            loss.backward()
            q = p.detach().clone().double().requires_grad_(True)          # float64 restatement of explainer.py:432-453
            x = q.clamp(1e-6, 1 - 1e-6)
            if prior == "empirical":
                s = x.mean(1, keepdim=True)
                oh = torch.nn.functional.one_hot(cat.long(), 12).double()
                m_c = (oh * x.unsqueeze(-1)).sum(1) / oh.sum(1).clamp(min=1)
                e = s * m_c
                nd = target * null.double().view(1, 12)
                ref = ((1 - s) * torch.log((1 - s) / (1 - target + 1e-6) + 1e-6) + e * torch.log(e / (nd + 1e-6) + 1e-6)).mean()
            else:
                ref = (x * torch.log(x / target + 1e-6) + (1 - x) * torch.log((1 - x) / (1 - target + 1e-6) + 1e-6)).mean()
            ref.backward()
            np.testing.assert_allclose(p.grad.cpu().numpy(), q.grad.float().cpu().numpy(), rtol=2e-4, atol=1e-7)


@pytest.mark.parametrize("M,N,K", [(128, 64, 32), (300, 172, 347), (27000, 64, 172), (1, 1, 64), (77, 76, 76), (257, 130, 9), (5, 300, 40)])
def test_tcgen05_gemm_of_the_training_path(tm, M, N, K):
    """tm_gemm_tf32x3: C = A B^T + bias (and the accumulate form) against float64; TcLinear's three products against autograd of F.linear."""
    from tempme_b200.training import TcLinear, gemm
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn((M, K), generator=g, device="cuda"); b = torch.randn((N, K), generator=g, device="cuda"); bias = torch.randn(N, generator=g, device="cuda")
    ref = (a.double() @ b.double().t() + bias.double())
    scale = (a.double().abs() @ b.double().abs().t()).clamp(min=1.0)
    out = gemm(a, b, bias)
    assert float(((out.double() - ref).abs() / scale).max()) < 2e-6           # 3xTF32: ~1e-6 of the sum of magnitudes (fp32: ~1e-7)
    out2 = gemm(a, b, None, out=out.clone(), accumulate=True)
    assert float(((out2.double() - (2 * ref - bias.double())).abs() / scale).max()) < 4e-6
    x = a.clone().requires_grad_(True); w = b.clone().requires_grad_(True); bb = bias.clone().requires_grad_(True)
    y = TcLinear.apply(x.view(1, M, K), w, bb)
    gy = torch.randn((1, M, N), generator=g, device="cuda")
    y.backward(gy)
    x64, w64, b64 = (t.detach().double().requires_grad_(True) for t in (a, b, bias))
    torch.nn.functional.linear(x64.view(1, M, K), w64, b64).backward(gy.double())
    for got, want in ((x.grad, x64.grad), (w.grad, w64.grad), (bb.grad, b64.grad)):
        s_ = max(float(want.abs().max()), 1e-12)
        assert float((got.double() - want).abs().max()) <= 3e-5 * s_ * max(1.0, (M * 1.0) ** 0.5 / 16), (M, N, K)
