"""GPU: SURVEY 8(f1) -- FusedAdam (spf_grad_sumsq + spf_adam_step) against the reference's optimiser step,
torch.nn.utils.clip_grad_norm_(1.0) -> NaN/Inf guard -> torch.optim.Adam.step (spurfies/train.py:355-363, 548-564),
run with plain torch ops on the same seeded tensors.  fp32: 1e-6 relative after 5 steps."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(1000, 64), (1000, 32), (256, 103), (256,), (3, 256), (3,), ()]  # odd sizes: segments get padded to 16 B


def _make(seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g).cuda() for s in SHAPES]


def _grads(step, scale):
    g = torch.Generator().manual_seed(100 + step)
    return [(torch.randn(s, generator=g) * scale).cuda() for s in SHAPES]


def _reference_step(params, opt, grads, max_norm=1.0):
    for p, g in zip(params, grads):
        p.grad = g.clone()
    torch.nn.utils.clip_grad_norm_(params, max_norm)
    ok = all(bool(torch.isfinite(p.grad).all()) for p in params)   # on_after_backward (train.py:548-564)
    if not ok:
        opt.zero_grad()
    opt.step()


@pytest.mark.parametrize("scale", [1e-3, 10.0])   # below / above the clip threshold
def test_fused_adam_matches_torch_adam(scale):
    from spurfies_b200.optim import FusedAdam
    ref_p = [torch.nn.Parameter(t.clone()) for t in _make(0)]
    our_p = [torch.nn.Parameter(t.clone()) for t in _make(0)]
    ref_opt = torch.optim.Adam(ref_p, lr=5e-4)
    ours = FusedAdam(our_p, lr=5e-4, max_norm=1.0)
    for step in range(5):
        grads = _grads(step, scale)
        if step == 2:
            grads[3][7] = float("nan")   # a poisoned step must be skipped entirely
        _reference_step(ref_p, ref_opt, grads)
        for p, g in zip(our_p, grads):
            p.grad.copy_(g)
        ours.step()
        assert bool(ours.skipped()) == (step == 2)
        assert float(ours.flat_g.abs().sum()) == 0.0   # zero_grad fused into the step
    for a, b in zip(our_p, ref_p):
        err = float((a.detach() - b.detach()).abs().max() / b.detach().abs().max().clamp(min=1e-12))
        assert err < 1e-6, err
    assert float(ours.state[0]) == 4.0   # 5 calls, one skipped
    # checkpoint layout of torch.optim.Adam: moments round-trip and match the reference optimiser's
    sd = ours.state_dict()
    rs = ref_opt.state_dict()["state"]
    for i in range(len(SHAPES)):
        for k in ("exp_avg", "exp_avg_sq"):
            want = rs[i][k]
            err = float((sd["state"][i][k] - want).abs().max() / want.abs().max().clamp(min=1e-20))
            assert err < 1e-5, (i, k, err)
    fresh = FusedAdam([torch.nn.Parameter(t.clone()) for t in _make(0)], lr=5e-4)
    fresh.load_state_dict(ref_opt.state_dict())
    assert float(fresh.state[0]) == 4.0
    assert torch.equal(fresh.view_of(fresh.exp_avg, 0), rs[0]["exp_avg"])


def test_grad_scale_is_the_data_parallel_average():
    """step(grad_scale = 1/W) on summed gradients == step on averaged gradients (norm and clip included)."""
    from spurfies_b200.optim import FusedAdam
    a = FusedAdam([torch.nn.Parameter(t.clone()) for t in _make(1)], lr=1e-3)
    b = FusedAdam([torch.nn.Parameter(t.clone()) for t in _make(1)], lr=1e-3)
    grads = _grads(0, 3.0)
    for p, g in zip(a.params, grads):
        p.grad.copy_(g * 4.0)
    for p, g in zip(b.params, grads):
        p.grad.copy_(g)
    a.step(grad_scale=0.25)
    b.step()
    assert abs(float(a.total_norm()) - float(b.total_norm())) < 1e-5 * float(b.total_norm())
    for x, y in zip(a.params, b.params):
        assert float((x.detach() - y.detach()).abs().max()) < 1e-7


def test_train_step_uses_fused_optimizer_and_learns():
    """TrainStep end to end (small scene): parameters are views of the flat buffer, the loss goes down."""
    from spurfies_b200 import scenes
    from spurfies_b200.model import PointVolSDF, default_conf
    from spurfies_b200.train import TrainStep
    sc = scenes.dtu_like(8000, seed=1, radii=(0.35, 0.5))
    torch.manual_seed(0)
    model = PointVolSDF(default_conf(), "24", "dtu", neural_points=sc["pts"], neural_colors=sc["colors"], precision="bf16")
    with torch.no_grad():
        model.neural_feats_geometry.mul_(8.0)
    step = TrainStep(model)
    base = step.opt.flat_p.data_ptr()
    assert all(base <= p.data_ptr() < base + 4 * step.opt.numel for p in step.params)
    R = 256
    cam = scenes.camera(0, sc["cam_radius"])
    batch = {"uv": scenes.pixel_batch(R, 3).cuda(), "pose": cam["pose"].cuda(), "intrinsics": cam["intrinsics"].cuda(),
             "local_data": None}
    gt = {k: v.cuda() for k, v in scenes.synthetic_gt(R, 3).items()}
    rng = {k: v.cuda() for k, v in scenes.rng_inputs(R, 3).items()}
    before = step.opt.flat_p.clone()
    losses = [float(step(batch, gt, rng)["loss"]) for _ in range(30)]
    assert all(l == l for l in losses)
    assert float(step.opt.state[0]) == 30.0 and float(step.opt.skipped()) == 0.0
    moved = (step.opt.flat_p - before).abs().max()
    assert 0.0 < float(moved) <= 30 * 5.0e-4 * 4.0       # Adam's per-step move is O(lr) whatever the gradient scale
    assert min(losses[10:]) < losses[0]                  # same batch every step: the loss goes down


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_direct_grad_accumulation_equals_autograd_accumulation(precision):
    """The scatter-add / split-K kernels writing straight into the flat gradient buffer (TrainStep's opt-in,
    fields._direct_grad) give the same gradients as zero-filled temporaries + autograd's AccumulateGrad."""
    from spurfies_b200 import scenes
    from spurfies_b200.model import PointVolSDF, VolSDFLoss, default_conf
    from spurfies_b200.optim import FusedAdam
    sc = scenes.dtu_like(8000, seed=1, radii=(0.35, 0.5))
    R = 256
    cam = scenes.camera(0, sc["cam_radius"])
    batch = {"uv": scenes.pixel_batch(R, 3).cuda(), "pose": cam["pose"].cuda(), "intrinsics": cam["intrinsics"].cuda(),
             "local_data": None}
    gt = {k: v.cuda() for k, v in scenes.synthetic_gt(R, 3).items()}
    rng = {k: v.cuda() for k, v in scenes.rng_inputs(R, 3).items()}
    flats = {}
    for direct in (False, True):
        torch.manual_seed(0)
        model = PointVolSDF(default_conf(), "24", "dtu", neural_points=sc["pts"], neural_colors=sc["colors"], precision=precision)
        with torch.no_grad():
            model.neural_feats_geometry.mul_(8.0)
        for prm in list(model.F_geometry.parameters()) + list(model.T.parameters()):
            prm.requires_grad_(False)
        opt = FusedAdam([p for p in model.parameters() if p.requires_grad])   # attaches every p.grad to one flat buffer
        opt.zero_grad()
        for p in opt.params:        # latent tables (both modes) and, in bf16 mode, the split-K weight-gradient outputs
            p._spf_direct_grad = direct
        model.train()
        out = model(batch, fast=1, rng=rng, dense_outputs=True)
        VolSDFLoss()(out, gt)["loss"].backward()
        assert opt.grads_attached()
        flats[direct] = opt.flat_g.clone()
    a, b = flats[True], flats[False]
    assert float(b.abs().max()) > 0
    assert float((a - b).abs().max()) <= (1e-5 if precision == "fp32" else 1e-4) * float(b.abs().max())   # atomics order only


def test_prefetched_inputs_give_the_same_step():
    """TrainStep.prefetch / step_prefetched (pinned host inputs copied on a side stream, eager and graph-replayed) run
    the same step as a direct call with device inputs: same first-step loss from the same initial weights."""
    from spurfies_b200 import scenes
    from spurfies_b200.model import PointVolSDF, default_conf
    from spurfies_b200.train import TrainStep
    sc = scenes.dtu_like(8000, seed=1, radii=(0.35, 0.5))
    R = 256
    cam = scenes.camera(0, sc["cam_radius"])
    host = ({"uv": scenes.pixel_batch(R, 3).pin_memory(), "pose": cam["pose"].pin_memory(),
             "intrinsics": cam["intrinsics"].pin_memory(), "local_data": None},
            {k: v.pin_memory() for k, v in scenes.synthetic_gt(R, 3).items()},
            {k: v.pin_memory() for k, v in scenes.rng_inputs(R, 3).items()})
    dev = tuple({k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()} for d in host)

    def fresh():
        torch.manual_seed(0)
        m = PointVolSDF(default_conf(), "24", "dtu", neural_points=sc["pts"], neural_colors=sc["colors"], precision="bf16")
        with torch.no_grad():
            m.neural_feats_geometry.mul_(8.0)
        return TrainStep(m, lr_schedule=False)

    ref = float(fresh()(*dev)["loss"])
    s1 = fresh()
    s1.prefetch(*host)
    eager = float(s1.step_prefetched()["loss"])
    s2 = fresh()
    assert s2.capture(*dev), s2.graph_error
    s3 = fresh()
    assert s3.capture(*dev)
    a = float(s2(*dev)["loss"])
    s3.prefetch(*host)
    b = float(s3.step_prefetched()["loss"])
    s3.prefetch(*host)
    c = float(s3.step_prefetched()["loss"])
    assert abs(eager - ref) <= 1e-4 * abs(ref), (eager, ref)
    # capture() restores parameters / moments / step count after its warm-up steps, so the first replayed step IS the
    # first step of training: same loss as the eager run from the same initial weights
    assert abs(a - ref) <= 1e-4 * abs(ref) and abs(b - ref) <= 1e-4 * abs(ref), (a, b, ref)
    assert c == c and c != b, (b, c)


def _fresh_step(sc, precision="bf16", **kw):
    from spurfies_b200.model import PointVolSDF, default_conf
    from spurfies_b200.train import TrainStep
    torch.manual_seed(0)
    m = PointVolSDF(default_conf(), "24", "dtu", neural_points=sc["pts"], neural_colors=sc["colors"], precision=precision)
    with torch.no_grad():
        m.neural_feats_geometry.mul_(8.0)
    return TrainStep(m, **kw)


def test_capture_has_no_side_effects_on_training_state():
    """capture()'s warm-up steps are real optimisation steps; everything they change (parameters, Adam
    moments, step count, lr, gradient buffer) must be restored so that graph training == eager training."""
    from spurfies_b200 import scenes
    sc = scenes.dtu_like(8000, seed=1, radii=(0.35, 0.5))
    R = 256
    cam = scenes.camera(0, sc["cam_radius"])
    batch = {"uv": scenes.pixel_batch(R, 3).cuda(), "pose": cam["pose"].cuda(), "intrinsics": cam["intrinsics"].cuda(),
             "local_data": None}
    gt = {k: v.cuda() for k, v in scenes.synthetic_gt(R, 3).items()}
    rng = {k: v.cuda() for k, v in scenes.rng_inputs(R, 3).items()}
    st = _fresh_step(sc)
    before = [t.clone() for t in (st.opt.flat_p, st.opt.exp_avg, st.opt.exp_avg_sq, st.opt.state, st.opt.flat_g)]
    assert st.capture(batch, gt, rng), st.graph_error
    after = (st.opt.flat_p, st.opt.exp_avg, st.opt.exp_avg_sq, st.opt.state, st.opt.flat_g)
    assert all(torch.equal(a, b) for a, b in zip(before, after))
    assert st.iter_step == 0 and float(st.opt.state[0]) == 0.0
    # three replayed steps == three eager steps (same batches), up to the atomics' summation order
    eager = _fresh_step(sc)
    for i in range(3):
        rg = {k: v.cuda() for k, v in scenes.rng_inputs(R, 3 + i).items()}
        la, lb = st(batch, gt, rg), eager(batch, gt, rg)
        assert abs(float(la["loss"]) - float(lb["loss"])) <= 2e-4 * abs(float(lb["loss"])), (i, float(la["loss"]), float(lb["loss"]))
    assert float(st.opt.state[0]) == 3.0 and st.iter_step == 3
    d = (st.opt.flat_p - eager.opt.flat_p).abs().max()
    assert float(d) <= 3 * 5e-4 * 0.5, float(d)
    # checkpoint round trip carries the scheduler position
    sd = st.state_dict()
    other = _fresh_step(sc)
    other.load_state_dict(sd)
    assert other.iter_step == 3 and float(other.opt.state[0]) == 3.0


def test_graph_replay_follows_changing_local_data():
    """the reference's main DTU config changes ``local_data`` (feature maps, cameras) every step; a captured
    step must read the NEW maps.  Graph replay vs eager over two steps with different local_data."""
    from spurfies_b200 import scenes
    sc = scenes.dtu_like(8000, seed=1, radii=(0.35, 0.5))
    R = 256
    ld = [{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in scenes.local_data(v_, sc["cam_radius"], feat_res=(128, 96), seed=v_).items()}
          for v_ in (0, 1)]

    def inputs(i):
        cam = scenes.camera(i, sc["cam_radius"])
        b = {"uv": scenes.pixel_batch(R, 3 + i).cuda(), "pose": cam["pose"].cuda(), "intrinsics": cam["intrinsics"].cuda(),
             "local_data": ld[i]}
        return (b, {k: v.cuda() for k, v in scenes.synthetic_gt(R, 3 + i).items()},
                {k: v.cuda() for k, v in scenes.rng_inputs(R, 3 + i).items()})

    graph, eager = _fresh_step(sc), _fresh_step(sc)
    assert graph.capture(*inputs(0)), graph.graph_error
    for i in (0, 1, 0):
        a, b = graph(*inputs(i)), eager(*inputs(i))
        for k in ("loss", "local_loss", "rgb_loss"):
            assert abs(float(a[k]) - float(b[k])) <= 2e-4 * max(abs(float(b[k])), 1e-3), (i, k, float(a[k]), float(b[k]))
    # a step whose inputs do not fit the captured structure is an error, never a silent replay of stale data
    bad = inputs(1)
    bad[0]["local_data"] = None
    with pytest.raises(ValueError):
        graph(*bad)


def test_arena_reuse_before_backward_is_an_error():
    """a second forward of the same model before the first backward overwrites the arena buffers holding
    the first one's saved activations -- that must raise, and two models must not share buffers at all."""
    from spurfies_b200 import scenes
    from spurfies_b200.model import PointVolSDF, VolSDFLoss, default_conf
    sc = scenes.dtu_like(8000, seed=1, radii=(0.35, 0.5))
    R = 128
    cam = scenes.camera(0, sc["cam_radius"])
    batch = {"uv": scenes.pixel_batch(R, 3).cuda(), "pose": cam["pose"].cuda(), "intrinsics": cam["intrinsics"].cuda(),
             "local_data": None}
    gt = {k: v.cuda() for k, v in scenes.synthetic_gt(R, 3).items()}
    rng = {k: v.cuda() for k, v in scenes.rng_inputs(R, 3).items()}
    torch.manual_seed(0)
    m1 = PointVolSDF(default_conf(), "24", "dtu", neural_points=sc["pts"], neural_colors=sc["colors"], precision="bf16")
    torch.manual_seed(1)
    m2 = PointVolSDF(default_conf(), "24", "dtu", neural_points=sc["pts"], neural_colors=sc["colors"], precision="bf16")
    for m in (m1, m2):
        m.train()
        with torch.no_grad():
            m.neural_feats_geometry.mul_(8.0)
    # (1) same model, forward twice, backward of the first: error
    l1 = VolSDFLoss()(m1(batch, fast=1, rng=rng, dense_outputs=True), gt)["loss"]
    _ = m1(batch, fast=1, rng=rng, dense_outputs=True)
    with pytest.raises(RuntimeError, match="overwritten by a later forward"):
        l1.backward()
    # (2) two models interleaved: each backward sees its own activations (== the gradients of a solo run)
    m1.zero_grad()
    VolSDFLoss()(m1(batch, fast=1, rng=rng, dense_outputs=True), gt)["loss"].backward()
    solo = m1.F_color[0].weight.grad.clone()
    m1.zero_grad()
    la = VolSDFLoss()(m1(batch, fast=1, rng=rng, dense_outputs=True), gt)["loss"]
    lb = VolSDFLoss()(m2(batch, fast=1, rng=rng, dense_outputs=True), gt)["loss"]
    la.backward()
    lb.backward()
    inter = m1.F_color[0].weight.grad
    assert float((inter - solo).abs().max()) <= 1e-4 * float(solo.abs().max())


@pytest.mark.parametrize("n,scale", [(1, 3.0e-7), (1000, 1.0), (983_040, 2.5e-4), (5_000_000, 7.0e3)])
def test_grad_scale_kernel_matches_torch_expression(n, scale):
    """spf_grad_scale == the torch expression it replaces (S = 2^floor(log2(target / max|x|)), [S, 1/S]); NaN in, NaN out;
    repeated calls (the scratch words are left zero)."""
    from spurfies_b200.fields import grad_scale
    g = torch.Generator().manual_seed(n)
    x = (torch.randn(n, generator=g) * scale).cuda()
    for target in (16.0, 1.0):
        amax = x.abs().amax().clamp(min=1.0e-30)
        e = torch.floor(torch.log2(target / amax)).clamp(-100.0, 100.0)
        want = torch.stack([torch.exp2(e), torch.exp2(-e)])
        for _ in range(2):
            got = grad_scale(x, target)
            assert torch.equal(got, want), (got, want)
    x[n // 2] = float("nan")
    assert bool(torch.isnan(grad_scale(x)).all())
    assert torch.equal(grad_scale(torch.zeros(8, device="cuda")), torch.tensor([2.0 ** 100, 2.0 ** -100], device="cuda"))
