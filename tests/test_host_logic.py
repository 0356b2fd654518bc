"""Host-side logic that needs no GPU: architecture tables, state_dict compatibility, launch-program
construction (trace-only), data-parallel helpers under gloo with world_size 2."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import hvae_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRESETS = ["morphomnist", "cmnist", "ukbb192", "mimic192", "mimic224", "tiny_ukbb"]


@pytest.mark.parametrize("name", PRESETS)
def test_state_dict_keys_and_shapes_match_reference_layout(name):
    from causalgen_b200 import HVAE
    cfg = O.make_cfg(name)
    model = HVAE(cfg)
    ours = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert list(ours.items()) == list(O.param_shapes(cfg).items())
    assert [b.res for b in model.decoder.blocks] == [d.res for d in O.build_arch(cfg).dec]
    assert [b.stochastic for b in model.decoder.blocks] == [d.stochastic for d in O.build_arch(cfg).dec]


@pytest.mark.parametrize("name", PRESETS)
def test_arch_tables_agree_with_oracle(name):
    from causalgen_b200.arch import decoder_plan, encoder_plan
    from causalgen_b200.presets import make_args
    a = make_args(name) if name != "tiny_ukbb" else make_args("tiny_ukbb")
    ref = O.build_arch(O.make_cfg(name))
    enc = encoder_plan(a)
    assert [(s.cin, s.cmid, s.cout, s.down) for s in enc] == [(b.cin, b.cmid, b.cout, b.down) for b in ref.enc]
    dec = decoder_plan(a)
    assert [(s.res, s.cin, s.cout, s.stochastic, s.ksize) for s in dec] == \
        [(d.res, d.cin, d.cout, d.stochastic, d.conv.ksize) for d in ref.dec]
    # odd resolutions are zero padded up by one after pooling (src/vae.py:130-132)
    if name == "mimic224":
        assert [s.res_out for s in enc if s.down][-2] == 8


def test_reference_init_scaling_rules():
    from causalgen_b200 import HVAE
    from causalgen_b200.presets import init_like_reference_main, make_args
    torch.manual_seed(7)
    m = init_like_reference_main(HVAE(make_args("tiny_ukbb")))
    sd = m.state_dict()
    assert float(sd["decoder.blocks.0.prior.conv.3.weight"].abs().max()) == 0.0  # src/vae.py:308
    assert all(float(v.abs().max()) == 0.0 for k, v in sd.items() if k.endswith(".bias") and "decoder.bias" not in k)
    assert float(sd["encoder.blocks.0.conv.1.weight"].abs().max()) > 0.0


def test_trace_only_program_construction():
    """builds the launch programs on CPU (nothing executes) and checks their structure"""
    code = r'''
import os, sys
os.environ["CAUSALGEN_B200_TRACE_ONLY"] = "1"
sys.path.insert(0, os.path.join(%r, "causal-gen_b200")); sys.path.insert(0, os.path.join(%r, "oracle"))
import torch, hvae_oracle as O
import causalgen_b200._lib as L
from causalgen_b200 import HVAE
class Fake:
    def __init__(self, real): self.real = real
    def __getattr__(self, n):
        if n in ("cg_conv_nchunk", "cg_conv_nchunk_ex", "cg_packed_weight_bytes", "cg_packed_weight_bytes_nc", "cg_version",
                     "cg_last_error"): return getattr(self.real, n)
        return lambda *a: 0
L._lib = Fake(L.load())
for name, nsto, nconv_min in (("tiny_ukbb", 5, 60), ("tiny_morphomnist", 4, 80)):
    cfg = O.make_cfg(name); m = HVAE(cfg)
    x8, pa, cf = O.synthetic_batch(cfg, 2, 1); x = O.normalise_x(x8)
    out = m(x, pa, beta=cfg.beta); out["elbo"].backward()
    prog = m.engine().programs[("elbo", 2, True, False)]
    names = [getattr(l, "name", "py") for l in prog.launches]
    assert names.count("cg_latent_fwd") == len(O.build_arch(cfg).dec)
    assert sum(n == "cg_conv2d" for n in names) >= nconv_min
    assert names.count("cg_conv2d_wgrad") >= 20 and names.count("cg_stem_wgrad") == 1
    assert prog.kl_rows.shape == (nsto, 2)
    fwd = names[:prog.n_fwd]
    assert "cg_conv2d_wgrad" not in fwd and fwd[-1] == "cg_elbo_finalize"
    zs = m.abduct(x, pa, t=0.9)
    assert len(zs) == nsto
    # kl_free_bits > 0 (src/vae.py:443-449): statistics -> floor / gate kernel -> finalize over ONE row; gated backward
    mf = HVAE(O.make_cfg(name, kl_free_bits=0.1))
    out = mf(x, pa, beta=cfg.beta); out["elbo"].backward()
    pf = mf.engine().programs[("elbo", 2, True, False)]
    nf = [getattr(l, "name", "py") for l in pf.launches]
    assert nf.count("cg_free_bits") == 1 and nf[pf.n_fwd - 1] == "cg_elbo_finalize" and nf[pf.n_fwd - 2] == "cg_free_bits"
    assert pf.kl_ch.shape == (nsto, 16) and pf.kl_gate.shape == (nsto, 16)
    assert sum(1 for lb in pf.D.latent_bwd_args if lb.kl_gate) == nsto and all(la.kl_ch for la in pf.D.latent_args if la.mode == 0)
    # lanes: every fork is joined again, pool (weight-gradient) launches never sit in the forward part
    kinds = [getattr(l, "kind", None) for l in prog.launches]
    assert kinds.count("fork") == kinds.count("join") and kinds.count("fork") >= nsto
    assert not any(getattr(l, "side", False) for l in prog.launches[:prog.n_fwd])
    assert all(l.side for l in prog.launches if getattr(l, "name", "") == "cg_conv2d_wgrad")
    # fused counterfactual program: encoder + posterior pass + two prior-only passes sharing the latents + combine
    from causalgen_b200 import counterfactual
    cf_prog = m.engine().build_counterfactual(2)
    cn = [getattr(l, "name", "py") for l in cf_prog.launches]
    nblk = len(O.build_arch(cfg).dec)
    assert cn.count("cg_stem_fwd") == 1 and cn[-1] == "cg_cf_combine" and "cg_nchw_f32_to_planar" not in cn
    assert cn.count("cg_latent_fwd") == nblk + 2 * (nblk - nsto)   # given latents need no latent kernel
    assert cn.count("cg_dgauss_sample") + cn.count("cg_dmol_predict") == 2
    print(name, len(names), prog.n_kernels, len(cn))
print("ok")
''' % (ROOT, ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "causal-gen_b200"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from causalgen_b200 import dp
    import hvae_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = O.make_cfg("tiny_ukbb")
    sd = {k: v.clone().requires_grad_(True) for k, v in O.seeded_state_dict(cfg, 7).items()}
    x8, pa, _ = O.synthetic_batch(cfg, 4, seed=3)
    x, pa_full = O.normalise_x(x8), O.expand_parents(pa, cfg.input_res)
    eps = O.NoiseTape(seed=5)
    with torch.no_grad():
        O.hvae_forward(sd, cfg, x, pa_full, eps)  # draw the global eps once (identical on both ranks)
    lo, hi = dp.shard_batch(4, world, rank)
    shard_eps = O.NoiseTape([e[lo:hi] for e in eps.drawn])
    out = O.hvae_forward(sd, cfg, x[lo:hi], pa_full[lo:hi], shard_eps, beta=cfg.beta)
    out["elbo"].backward()
    flat = torch.cat([p.grad.reshape(-1) for p in sd.values() if p.grad is not None])
    params = torch.cat([p.detach().reshape(-1) for p in sd.values()]) + (rank * 1.0)
    dp.broadcast_params_(params)
    scale = dp.reduce_gradients_(flat)
    q.put((rank, (flat * scale).numpy(), params.numpy(), dp.rank_noise_seed(7, rank)))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_gradient_equals_full_batch_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    g0, g1 = res[0][1], res[1][1]
    np.testing.assert_array_equal(g0, g1)  # every rank holds the same reduced bucket
    np.testing.assert_array_equal(res[0][2], res[1][2])  # broadcast made the parameters identical
    assert res[0][3] != res[1][3]  # rank-disjoint noise streams
    # single-process full batch
    cfg = O.make_cfg("tiny_ukbb")
    sd = {k: v.clone().requires_grad_(True) for k, v in O.seeded_state_dict(cfg, 7).items()}
    x8, pa, _ = O.synthetic_batch(cfg, 4, seed=3)
    out = O.hvae_forward(sd, cfg, O.normalise_x(x8), O.expand_parents(pa, cfg.input_res), O.NoiseTape(seed=5), beta=cfg.beta)
    out["elbo"].backward()
    full = torch.cat([p.grad.reshape(-1) for p in sd.values() if p.grad is not None]).numpy()
    np.testing.assert_allclose(g0, full, rtol=2e-4, atol=1e-6)


def test_shard_batch_rejects_ragged_split():
    from causalgen_b200 import dp
    assert dp.shard_batch(32, 4, 3) == (24, 32)
    with pytest.raises(ValueError):
        dp.shard_batch(30, 4, 0)


def _happens_before(plan, n_launches):
    """replay a stream plan: for every launch, the set of launches guaranteed complete before it starts"""
    done = {}      # stream -> launches complete before anything issued later on that stream
    issued = {}    # stream -> launches issued on that stream so far
    snap = {}      # event id -> launches covered by the event
    before = {}
    for act in plan:
        kind, st, x = act
        done.setdefault(st, set())
        issued.setdefault(st, set())
        if kind == "record":
            snap[x] = done[st] | issued[st]
        elif kind == "wait":
            done[st] |= snap[x]
        else:
            before[x] = done[st] | issued[st]
            issued[st] = issued[st] | {x}
    assert sorted(before) == list(range(n_launches))
    return before, done.get(0, set()) | issued.get(0, set())


def test_stream_plan_respects_fork_join_and_side_dependencies():
    """the pure scheduler behind Program.run: lane order, fork/join edges, pool launches waiting on their lane --
    including the case of a pool launch recorded on the auxiliary lane right after a fork (it must inherit the
    main-stream work the fork waited for)"""
    from types import SimpleNamespace as NS
    from causalgen_b200.engine import Marker, plan_streams

    def L(lane=0, side=False):
        return NS(lane=lane, side=side)

    # index:      0        1             2                 3           4              5        6
    prog = [L(0), L(0), Marker("fork"), L(1, side=True), L(1), Marker("join"), L(0), L(0, side=True), L(0)]
    launches = [i for i, ln in enumerate(prog) if not isinstance(ln, Marker)]
    for n_sides in (1, 2, 3, 6):   # 6 = the default pool width (engine.SIDE_STREAMS)
        plan = plan_streams(prog, n_sides)
        done_sets, final = _happens_before([a if a[0] != "launch" else ("launch", a[1], launches.index(a[2])) for a in plan],
                                          len(launches))
        idx = {orig: launches.index(orig) for orig in launches}
        assert {idx[0], idx[1]} <= done_sets[idx[3]], "pool launch on the aux lane must wait for pre-fork main work"
        assert {idx[0], idx[1]} <= done_sets[idx[4]], "aux lane starts after the fork point"
        assert idx[4] in done_sets[idx[6]], "main continues only after the join"
        assert {idx[0], idx[1], idx[4], idx[6]} <= done_sets[idx[7]], "pool launch on main waits for main so far"
        assert idx[6] in done_sets[idx[8]]
        assert final == set(range(len(launches))), "everything is joined back into the main stream at the end"
    # second block: the auxiliary lane's cached event predates the next fork -- a pool launch recorded right after
    # that fork must still see the main-stream work issued in between (latent backward -> posterior weight gradient)
    prog = [Marker("fork"), L(1), Marker("join"), L(0), Marker("fork"), L(1, side=True), L(1), Marker("join"), L(0)]
    launches = [i for i, ln in enumerate(prog) if not isinstance(ln, Marker)]
    for n_sides in (1, 2, 6):
        plan = plan_streams(prog, n_sides)
        done_sets, final = _happens_before([a if a[0] != "launch" else ("launch", a[1], launches.index(a[2])) for a in plan],
                                          len(launches))
        a1, m, w, a2, tail = range(5)
        assert {a1, m} <= done_sets[w], "pool launch after the second fork must wait for the main work before it"
        assert {a1, m} <= done_sets[a2] and a2 in done_sets[tail]
        assert final == set(range(5))


def test_launch_policies_on_the_recorded_programs():
    """the two scheduling policies as they land in the launch arguments of real programs (built on the CPU, nothing runs):
    column-folded convs exactly on the layers ops.fold_pays names, cg_wgrad_args.min_tiles from ops.wgrad_min_tiles"""
    code = r"""
import os, sys
os.environ["CAUSALGEN_B200_TRACE_ONLY"] = "1"
sys.path.insert(0, os.path.join(%r, "causal-gen_b200")); sys.path.insert(0, os.path.join(%r, "oracle"))
import torch, hvae_oracle as O
import causalgen_b200._lib as L
from causalgen_b200 import HVAE
class Fake:
    def __init__(self, real): self.real = real
    def __getattr__(self, n):
        if n in ("cg_conv_nchunk", "cg_conv_nchunk_ex", "cg_packed_weight_bytes", "cg_packed_weight_bytes_nc", "cg_version",
                 "cg_last_error", "cg_conv_fold_ok", "cg_conv2d_wgrad_launches"): return getattr(self.real, n)
        return lambda *a: 0
L._lib = Fake(L.load())
for name, light in (("ukbb192", True), ("morphomnist", False)):
    cfg = O.make_cfg(name); m = HVAE(cfg)
    prog = m.engine().build_elbo(1, True, False)
    folded, plain = set(), set()
    for ln in prog.launches:
        nm = getattr(ln, "name", "")
        if nm == "cg_conv2d":
            a = ln.keep[0]
            K = sum(a.src[s].C for s in range(a.nsrc))
            if a.ksize == 3 and a.cout <= 32 and K >= 2 * a.cout:
                (folded if a.fold else plain).add((a.H, K >= 64))
            else:
                assert a.fold == 0
        elif nm == "cg_conv2d_wgrad":
            assert ln.keep[0].min_tiles == (24 if light else 48), (name, ln.keep[0].min_tiles)   # 17*sqrt(1) clamps to 24
    if light:
        assert folded == {(192, True), (96, True), (24, True)}, folded      # wide K where the tile count stays within 20 %%
        assert (48, True) in plain and (192, False) in plain, plain          # 48^2: 24 against 18 tiles; 32 -> 8: two K-blocks
    else:
        assert not folded or all(k for _, k in folded)
print("ok")
""" % (ROOT, ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("name", ["tiny_ukbb", "tiny_morphomnist", "tiny_cmnist", "morphomnist", "cmnist", "ukbb192",
                                  "mimic192", "mimic224"])
def test_fresh_init_is_bit_identical_to_the_reference(name):
    """same seed -> same weights as the reference's HVAE(args): module construction order, default Conv2d init and the
    init scaling rules (src/vae.py:121-122,303-308), digests from tests/golden/make_golden_init.py"""
    import hashlib
    import json
    from causalgen_b200 import HVAE
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "init_digests.json")))[name]
    torch.manual_seed(7)
    model = HVAE(O.make_cfg(name))
    sd = model.state_dict()
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    assert len(sd) == want["tensors"] and sum(p.numel() for p in model.parameters()) == want["params"]
    assert h.hexdigest() == want["sha256"]


def test_vae_preprocess_matches_reference_functions():
    """src/pgm/dscm.py:98-132 incl. the UKBB log-standardisation branch, against vectors produced by executing the
    reference's own function bodies (tests/golden/make_golden_preprocess.py)"""
    from types import SimpleNamespace
    from causalgen_b200 import vae_preprocess
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess.npz"))
    pa = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("in_")}
    a = SimpleNamespace(dataset="ukbb", input_res=8, parents_x=["mri_seq", "brain_volume", "ventricle_volume", "sex"])
    src = {k: pa[k].clone() for k in a.parents_x}
    got = vae_preprocess(a, src).cpu()
    np.testing.assert_allclose(got.numpy(), g["ukbb_vae_pa"][:, :, 0, 0], rtol=1e-6, atol=1e-6)
    assert all(torch.equal(src[k], pa[k]) for k in src), "caller's dict must not be modified"
    full = vae_preprocess(a, {k: pa[k].clone() for k in a.parents_x}, expand=True).cpu()
    np.testing.assert_allclose(full.numpy(), g["ukbb_vae_pa"], rtol=1e-6, atol=1e-6)
    a2 = SimpleNamespace(dataset="ukbb", input_res=4, parents_x=["age", "sex"])
    got = vae_preprocess(a2, {k: pa[k].clone() for k in a2.parents_x}).cpu()
    np.testing.assert_allclose(got.numpy(), g["ukbb_age_vae_pa"][:, :, 0, 0], rtol=1e-6, atol=1e-6)
    a3 = SimpleNamespace(dataset="morphomnist", input_res=4, parents_x=["brain_volume", "digit"])
    got = vae_preprocess(a3, {"brain_volume": pa["brain_volume"][:, 0].clone(), "digit": pa["digit"]}).cpu()
    np.testing.assert_allclose(got.numpy(), g["plain_vae_pa"][:, :, 0, 0], rtol=0, atol=0)
    with pytest.raises(KeyError):
        vae_preprocess(a, {"mri_seq": pa["mri_seq"], "brain_volume": pa["brain_volume"], "ventricle_volume": pa["sex"],
                           "sex": pa["sex"], "thickness": pa["age"]})


def test_checkpoint_wire_format_round_trips_through_torch_and_the_reference():
    """src/trainer.py:154-168 / src/main.py:74-90: our checkpoint loads into real torch AdamW / LambdaLR objects (what the
    reference does on resume), a checkpoint written from the REFERENCE's own model + optimizer loads into a Trainer, and
    the file survives torch.save / torch.load.  Runs trace-only (no kernels) in a subprocess."""
    code = r'''
import os, sys, tempfile
os.environ["CAUSALGEN_B200_TRACE_ONLY"] = "1"
sys.path.insert(0, os.path.join(%r, "causal-gen_b200")); sys.path.insert(0, os.path.join(%r, "oracle"))
import torch, hvae_oracle as O
import causalgen_b200._lib as L
class Fake:
    def __init__(self, real): self.real = real
    def __getattr__(self, n):
        if n in ("cg_conv_nchunk", "cg_conv_nchunk_ex", "cg_packed_weight_bytes", "cg_packed_weight_bytes_nc", "cg_version",
                 "cg_last_error"): return getattr(self.real, n)
        return lambda *a: 0
L._lib = Fake(L.load())
from causalgen_b200 import HVAE, checkpoint as CK
from causalgen_b200.trainer import Trainer
from causalgen_b200.presets import make_args
cfg = make_args("morphomnist")
tr = Trainer(HVAE(cfg), 2, lr=1e-3, wd=0.01, lr_warmup_steps=100, use_graph=False)
g = torch.Generator().manual_seed(0)
tr.m.copy_(torch.randn(tr.m.shape, generator=g)); tr.v.copy_(torch.rand(tr.v.shape, generator=g))
tr.ema.copy_(torch.randn(tr.ema.shape, generator=g)); tr.state[0] = 37; tr.state[1] = 37; tr.steps_done = 37
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "checkpoint.pt")
    CK.save(path, tr, hparams={"lr": 1e-3, "hps": "morphomnist"}, epoch=3, best_loss=1.25)
    ck = torch.load(path, weights_only=False)
assert set(ck) >= {"epoch", "step", "best_loss", "model_state_dict", "ema_model_state_dict", "optimizer_state_dict",
                   "scheduler_state_dict", "hparams"}
assert list(ck["model_state_dict"]) == list(ck["ema_model_state_dict"]) == [k for k, _ in tr.model.named_parameters()]
# the reference's resume path: fresh model / AdamW / LambdaLR, load_state_dict of each (src/main.py:74-79)
m2 = HVAE(cfg); m2.load_state_dict(ck["model_state_dict"])
opt = torch.optim.AdamW(m2.parameters(), lr=1e-3, weight_decay=0.01, betas=(0.9, 0.9))
opt.load_state_dict(ck["optimizer_state_dict"])
assert abs(opt.param_groups[0]["lr"] - 0.37e-3) < 1e-12      # lr of optimizer.step() number 38 under the warm-up
sch = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda it: 1.0 if it > 100 else it / 100)
sch.load_state_dict(ck["scheduler_state_dict"])
assert sch.last_epoch == 37 and abs(sch.get_last_lr()[0] - 0.37e-3) < 1e-12
names = [k for k, _ in m2.named_parameters()]
mv = tr._views(tr.m)
for i, p in enumerate(m2.parameters()):
    if names[i] in mv:
        assert torch.equal(opt.state[p]["exp_avg"], mv[names[i]]) and float(opt.state[p]["step"]) == 37.0
    else:
        assert p not in opt.state          # last block's z_feat_proj: never used, no Adam state in the reference either
# and back into a fresh trainer
tr2 = Trainer(HVAE(cfg), 2, lr=1e-3, wd=0.01, lr_warmup_steps=100, use_graph=False)
CK.from_checkpoint(tr2, ck)
assert torch.equal(tr2.m, tr.m) and torch.equal(tr2.v, tr.v) and torch.equal(tr2.ema, tr.ema) and torch.equal(tr2.flat_p, tr.flat_p)
assert tr2.state.tolist() == [37, 37, 0, 0] and tr2.steps_done == 37
CK.from_checkpoint(tr2, ck, mode="reference")   # src/main.py:80-86: constant lr, fresh EMA counter
assert tr2.state.tolist() == [37, 0, 0, 0] and tr2.hp["warmup"] == 0
# a checkpoint produced by the reference's own objects (staged reference, when present)
sys.path.insert(0, os.path.join(%r, "oracle"))
import ref_runner as R
if R.available():
    st = R.RefTrainStep("morphomnist", "cpu")
    x8, pa, _ = O.synthetic_batch(O.make_cfg("morphomnist"), 2, seed=1)
    for _ in range(2): st(x8, pa)
    ref_ck = {"epoch": 1, "step": 2, "best_loss": 3.0, "model_state_dict": st.model.state_dict(),
              "ema_model_state_dict": st.ema.ema_model.state_dict(), "optimizer_state_dict": st.opt.state_dict(),
              "scheduler_state_dict": st.sched.state_dict(), "hparams": {"lr": 1e-3}}
    tr3 = Trainer(HVAE(cfg), 2, lr=1e-3, wd=0.01, lr_warmup_steps=100, use_graph=False)
    CK.from_checkpoint(tr3, ref_ck)
    m3 = tr3._views(tr3.m)
    ref_named = dict(st.model.named_parameters())
    n_checked = 0
    for k, p in ref_named.items():
        if p in st.opt.state:
            assert torch.equal(m3[k], st.opt.state[p]["exp_avg"]), k
            n_checked += 1
    assert n_checked > 100 and tr3.state.tolist()[:2] == [2, 2]
    assert all(torch.equal(dict(tr3.model.named_parameters())[k], v) for k, v in st.model.state_dict().items())
    print("REF_CKPT_OK")
print("CKPT_OK")
''' % (ROOT, ROOT, ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "CKPT_OK" in r.stdout, r.stderr[-3000:]
