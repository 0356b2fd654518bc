"""Launch-program builder for the HVAE hot path.

For a given (model, batch size, mode) the engine allocates every activation / gradient buffer once and
records the exact sequence of C-ABI kernel launches of the pass (`Program`).  Running a pass is then a
flat loop over pre-built argument structs -- no per-step allocation, CUDA-graph capturable.

Reference semantics implemented here (citations relative to /root/reference):
  encoder           src/vae.py:125-134        decoder loop      src/vae.py:222-301
  Block             src/vae.py:73-84          ELBO reduction    src/vae.py:439-458
  abduct / mixture  src/vae.py:466-516        forward_latents   src/vae.py:518-522
Backward passes are hand-derived (there is no autograd under this package).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib as L
from .model import Block, DmolNet
from .ops import ConvLayer, PackTable, SegSpec, View, new_act, round16


# CAUSALGEN_B200_TRACE_ONLY=1 lets the launch programs be *built* (buffers + argument structs) on a host
# without a GPU so their structure can be unit-tested; nothing is executed and outputs stay zero.
TRACE_ONLY = os.environ.get("CAUSALGEN_B200_TRACE_ONLY", "0") == "1"
# Weight-gradient launches only feed the flat gradient bucket, so they are forked onto side streams (graph
# branches under capture) and overlap the latency-bound data-gradient chain.  0 = everything on one stream.
SIDE_STREAMS = int(os.environ.get("CAUSALGEN_B200_SIDE_STREAMS", "6"))
# The two lanes carry the dependent chain of the step (every conv waits for its predecessor); the weight gradients on
# the pool streams are filler.  The lanes run at high stream priority (the capture stream of `Trainer` and the auxiliary
# lane), so a lane kernel's CTAs are placed before queued weight-gradient CTAs; the deferred weight gradients then need a
# wider pool to fill the gaps.  Measured together (profiles/r4b_streams_priority_ab.txt): priority with 2 pool streams
# loses 1 % (batch 128) / 5 % (batch 32), with 6 it is neutral at batch 128 and +3 % at batch 32.
# CAUSALGEN_B200_PRIO=0 restores equal priorities.
LANE_PRIORITY = -1 if os.environ.get("CAUSALGEN_B200_PRIO", "1") == "1" else 0


def lane_stream() -> "torch.cuda.Stream":
    """a stream for one of the two lanes (capture stream of the Trainer graphs, auxiliary lane)"""
    return torch.cuda.Stream(priority=LANE_PRIORITY)


class PyOp:
    """a recorded torch-side glue op (layout copies only; never arithmetic on the path)"""

    side = False
    lane = 0

    def __init__(self, fn, name="pyop"):
        self.fn, self.name = fn, name

    def __call__(self, stream):
        if self.lane == 0:
            self.fn()
        else:  # torch op recorded on the auxiliary lane: run it on that stream
            with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
                self.fn()


class Marker:
    """fork / join point between the main stream (lane 0) and the auxiliary lane of a program"""
    side = False
    lane = 0

    def __init__(self, kind):
        self.kind, self.name = kind, "lane_" + kind

    def __call__(self, stream):  # sequential execution: nothing to do
        pass


class HostOp:
    """A host-side step inside a program (a torch.distributed collective between two kernels): issued in order on the
    main lane like a launch.  Programs that hold one are not captured into CUDA graphs."""
    side = False
    name = "host_op"

    def __init__(self, what, fn):
        self.what, self.fn, self.lane, self.algo_bytes = what, fn, 0, 0

    def __call__(self, stream):
        self.fn(stream)


class Program:
    """Recorded launch sequence.  Dependencies are expressed structurally: launches on lane 0 (main stream) and on
    lane 1 (auxiliary stream) are each in order; ``fork()`` makes lane 1 wait for everything issued on lane 0 so far,
    ``join()`` the reverse.  Weight-gradient launches (``side``) only feed the gradient bucket: each waits for the
    lane it was recorded on and runs on a pool stream; everything meets again at the end of ``run``."""

    def __init__(self, name):
        self.name = name
        self.launches: List = []
        self.keep: List = []
        self.n_kernels = 0
        self.sides = None
        self.aux = None
        self.cur_lane = 0
        self._plan = None

    def fork(self):
        self.launches.append(Marker("fork"))

    def join(self):
        self.launches.append(Marker("join"))

    class _Lane:
        def __init__(self, prog, lane):
            self.prog, self.lane = prog, lane

        def __enter__(self):
            self.prev, self.prog.cur_lane = self.prog.cur_lane, self.lane

        def __exit__(self, *exc):
            self.prog.cur_lane = self.prev

    def on_lane(self, lane):
        return Program._Lane(self, lane)

    def add(self, ln):
        self.launches.append(ln)
        ln.lane = self.cur_lane
        if isinstance(ln, L.Launch):
            if ln.name == "cg_conv2d_wgrad":
                self.n_kernels += int(L.load().cg_conv2d_wgrad_launches(C.byref(ln.keep[0])))
                ln.side = True
            else:
                self.n_kernels += 1
        return ln

    def call(self, name, *args):
        return self.add(L.Launch(name, *args))

    def run(self, stream=None):
        if TRACE_ONLY:  # program construction check on a box without a GPU: nothing is computed
            return
        main = torch.cuda.current_stream()
        s = main.cuda_stream if stream is None else stream
        if SIDE_STREAMS <= 0 or s != main.cuda_stream:
            for ln in self.launches:
                ln(s)
            return
        if self.sides is None:
            self.sides = [torch.cuda.Stream() for _ in range(SIDE_STREAMS)]
            self.aux = lane_stream()
        streams = [main, self.aux] + self.sides          # stream ids of plan_streams(): 0 main, 1 aux, 2.. pool
        raw = [s, self.aux.cuda_stream] + [st.cuda_stream for st in self.sides]
        events = {}
        if self._plan is None or self._plan[0] != len(self.launches):
            self._plan = (len(self.launches), plan_streams(self.launches, len(self.sides)))
        for act in self._plan[1]:
            if act[0] == "record":
                events[act[2]] = torch.cuda.Event()
                events[act[2]].record(streams[act[1]])
            elif act[0] == "wait":
                streams[act[1]].wait_event(events[act[2]])
            else:
                self.launches[act[2]](raw[act[1]])


def plan_streams(launches, n_sides):
    """Pure scheduling of a recorded program onto streams (0 = main, 1 = auxiliary lane, 2.. = weight-gradient pool).
    Yields ("record", stream, event_id) / ("wait", stream, event_id) / ("launch", stream, launch_index) in issue
    order; ends with every stream joined back into main.  Kept free of CUDA so the dependency logic is unit-tested
    on the CPU (tests/test_host_logic.py)."""
    ev = [None, None]        # per lane: id of an event covering everything issued on it so far (None = stale)
    seen = [[None, None] for _ in range(n_sides)]
    next_id = [0]
    out = []

    def event_of(lane):
        if ev[lane] is None:
            ev[lane] = next_id[0]
            next_id[0] += 1
            out.append(("record", lane, ev[lane]))
        return ev[lane]

    k, aux_used = 0, False
    for idx, ln in enumerate(launches):
        if isinstance(ln, Marker):
            if ln.kind == "fork":
                out.append(("wait", 1, event_of(0)))
                ev[1] = None  # the lane now also covers main's work: a cached older event would miss it
                aux_used = True
            else:
                out.append(("wait", 0, event_of(1)))
                ev[0] = None
            continue
        if ln.side:
            j = k % n_sides
            k += 1
            e = event_of(ln.lane)
            if seen[j][ln.lane] != e:
                out.append(("wait", 2 + j, e))
                seen[j][ln.lane] = e
            out.append(("launch", 2 + j, idx))
        else:
            out.append(("launch", ln.lane, idx))
            ev[ln.lane] = None
    tail = ([1] if aux_used else []) + ([2 + j for j in range(n_sides)] if k else [])
    for st in tail:
        out.append(("record", st, next_id[0]))
        out.append(("wait", 0, next_id[0]))
        next_id[0] += 1
    return out


class BlockLayers:
    """ConvLayers of one reference Block"""

    def __init__(self, eng: "Engine", mod: Block, src_logical: Sequence[int], res: int, grad_srcs=None):
        self.mod = mod
        self.act = L.ACT_RELU if mod.light else L.ACT_GELU
        # ReLU commutes with bf16 rounding and relu(x) > 0 <=> x > 0, so the intermediate of a "light" block
        # is stored already activated by the producing conv: its consumers (next conv, weight gradient) skip
        # the pre-activation pass and the data-gradient mask reads the same tensor.
        self.mid_out_act = L.ACT_RELU if mod.light else L.ACT_NONE
        # GELU'(y) needs the raw y, so a GELU block stores BOTH: y (for the data-gradient factor) and gelu(y) (input of
        # the next conv and of its weight gradient) -- the intermediates are 4x narrower than the block, the second
        # store is cheap and removes every in-kernel GELU pass over them.
        self.mid_dual = not mod.light
        centre = (res == 1 and mod.ksize == 3)
        convs = mod.convs
        self.layers: List[ConvLayer] = []
        for i, c in enumerate(convs):
            srcs = list(src_logical) if i == 0 else [c.weight.shape[1]]
            act_i = L.ACT_NONE if (i > 0 and (self.mid_out_act != L.ACT_NONE or self.mid_dual)) else self.act
            self.layers.append(ConvLayer(eng.table, c.weight, c.bias, srcs, act_i,
                                         centre_only=centre and c.weight.shape[2] == 3,
                                         grad_srcs=(grad_srcs if i == 0 else None),
                                         fwd_operands=(i == len(convs) - 1),  # only a Block's last conv fuses adds
                                         res=res, light=mod.light))
        self.proj = None
        if hasattr(mod, "width_proj"):
            self.proj = ConvLayer(eng.table, mod.width_proj.weight, mod.width_proj.bias, list(src_logical), L.ACT_NONE,
                                  light=mod.light)
        self.params = [(c.weight, c.bias) for c in convs]


class Rec:
    """attribute bag of tensors a forward pass saves for its backward"""
    pass


def ordered_params(model) -> List[torch.nn.Parameter]:
    """parameter order of the flat buffers: trainable first, frozen last (so optimiser kernels can stop before the
    frozen tail: AdamW skips parameters without a gradient, src/train_setup.py:42-45 + src/vae.py:340-349)"""
    ps = list(model.parameters())
    dead = no_grad_param_ids(model)
    live = [p for p in ps if p.requires_grad and id(p) not in dead]
    return live + [p for p in ps if not (p.requires_grad and id(p) not in dead)]


def no_grad_param_ids(model):
    """parameters the forward never touches: the LAST decoder block's z_feat_proj (src/vae.py:297-300 skips it).  In the
    reference their .grad stays None, so AdamW neither decays nor updates them; they are treated like frozen ones."""
    blocks = getattr(getattr(model, "decoder", None), "blocks", None)
    if not blocks or not hasattr(blocks[-1], "z_feat_proj"):
        return set()
    return {id(p) for p in blocks[-1].z_feat_proj.parameters()}


class Engine:
    def __init__(self, model, args):
        L.load()
        self.model = model
        self.args = args
        p0 = next(model.parameters())
        if p0.device.type != "cuda" and not TRACE_ONLY:
            raise RuntimeError("causalgen_b200 runs on a CUDA sm_100 device only (no CPU fallback); move the model "
                               "to cuda first")
        self.device = p0.device
        if not TRACE_ONLY:
            with torch.cuda.device(self.device):
                if L.load().cg_device_sms() <= 0:
                    raise RuntimeError("causalgen_b200: device is not sm_100 (B200); there is no fallback path")
        self.table = PackTable(self.device)
        self.zd = args.z_dim
        self.ctx = args.context_dim
        self.ctx_pad = round16(self.ctx)
        self.cond_prior = bool(args.cond_prior)
        self.q_corr = bool(args.q_correction)
        self.C = args.input_channels
        self.R = args.input_res
        self.light = args.vr == "light"
        enc, dec = model.encoder, model.decoder
        self.enc_layers = [BlockLayers(self, b, [st.cin], st.res_in) for b, st in zip(enc.blocks, enc.plan)]
        self.dec_layers = []
        for blk, st in zip(dec.blocks, dec.plan):
            d = Rec()
            d.st = st
            psrc = [st.cin] + ([self.ctx] if self.cond_prior else [])
            d.prior = BlockLayers(self, blk.prior, psrc, st.res, grad_srcs=[True] + [False] * (len(psrc) - 1))
            d.post = None
            if st.stochastic:
                d.post = BlockLayers(self, blk.posterior, [st.cin, self.ctx, st.cin], st.res,
                                     grad_srcs=[True, False, True])
            d.z_proj = ConvLayer(self.table, blk.z_proj.weight, blk.z_proj.bias, [self.zd, self.ctx], L.ACT_NONE,
                                 grad_srcs=[True, False], light=blk.conv.light)
            d.zfp = None
            if not self.q_corr:
                d.zfp = ConvLayer(self.table, blk.z_feat_proj.weight, blk.z_feat_proj.bias, [self.zd, st.cin],
                                  L.ACT_NONE, light=blk.conv.light)
            d.conv = BlockLayers(self, blk.conv, [st.cin], st.res)
            self.dec_layers.append(d)
        self.dmol = isinstance(model.likelihood, DmolNet)
        # flat gradient bucket with per-parameter views (the one buffer DDP all-reduces)
        self.params = ordered_params(model)
        self.generation = 0  # bumped by every training forward: guards flat_grad against interleaved fwd/bwd pairs
        n = sum(p.numel() for p in self.params)
        self.flat_grad = torch.zeros(n, device=self.device, dtype=torch.float32)
        self.grad_of: Dict[int, torch.Tensor] = {}
        off = 0
        for p in self.params:
            self.grad_of[id(p)] = self.flat_grad[off: off + p.numel()].view_as(p)
            off += p.numel()
        self.signature = self.param_signature(model)
        self.programs: Dict = {}

    @staticmethod
    def param_signature(model):
        return tuple((p.data_ptr(), p.requires_grad) for p in model.parameters())

    def g(self, p: Optional[torch.Tensor]):
        return None if p is None else self.grad_of[id(p)]

    def pack_weights(self, stream=None):
        if TRACE_ONLY:
            return
        s = torch.cuda.current_stream().cuda_stream if stream is None else stream
        self.table.launch(s)

    # ================================================================== forward emission
    def _block_fwd(self, prog: Program, bl: BlockLayers, srcs: List[View], N, H, W, final_segs=None) -> Rec:
        """reference Block.forward (src/vae.py:73-84) without the pooling"""
        r = Rec()
        r.bl, r.srcs, r.N, r.H, r.W = bl, srcs, N, H, W
        r.mids = []       # what the data gradient multiplies by act'(.)
        r.mids_in = []    # what the next conv / its weight gradient consume (already activated when possible)
        cur = srcs
        for layer in bl.layers[:-1]:
            mid = new_act(N, H, W, layer.cout_l, self.device)
            if bl.mid_dual:
                mid_a = new_act(N, H, W, layer.cout_l, self.device)
                prog.add(layer.forward(cur, [SegSpec(mid, 0, out_act=bl.act, act_copy=mid_a)], N, H, W))
            else:
                mid_a = mid
                prog.add(layer.forward(cur, [SegSpec(mid, 0, out_act=bl.mid_out_act)], N, H, W))
            r.mids.append(mid)
            r.mids_in.append(mid_a)
            cur = [mid_a]
        last = bl.layers[-1]
        r.proj_out = None
        if final_segs is None:
            skip = None
            if bl.mod.residual:
                if bl.proj is not None:
                    r.proj_out = new_act(N, H, W, bl.proj.cout_l, self.device)
                    prog.add(bl.proj.forward(srcs, [SegSpec(r.proj_out, 0)], N, H, W))
                    skip = r.proj_out
                else:
                    skip = srcs[0]
            r.y = new_act(N, H, W, last.cout_l, self.device)
            final_segs = [SegSpec(r.y, 0, add=skip)]
        prog.add(last.forward(cur, final_segs, N, H, W))
        return r

    def _encoder_fwd(self, prog: Program, x: torch.Tensor, N) -> Rec:
        enc = self.model.encoder
        e = Rec()
        e.x = x
        R = self.R
        w0 = self.args.widths[0]
        e.stem_out = new_act(N, R, R, w0, self.device)
        prog.call("cg_stem_fwd", x.data_ptr(), enc.stem.weight.data_ptr(), enc.stem.bias.data_ptr(),
                  e.stem_out.ptr, N, self.C, R, w0, e.stem_out.ns)
        prog.keep.append(x)
        cur = e.stem_out
        e.blocks = []
        e.acts: Dict[int, View] = {}
        e.owner: Dict[int, int] = {}
        for i, (bl, st) in enumerate(zip(self.enc_layers, enc.plan)):
            r = self._block_fwd(prog, bl, [cur], N, st.res_in, st.res_in)
            r.st = st
            if st.down:
                r.out = new_act(N, st.res_out, st.res_out, st.cout, self.device)
                prog.call("cg_avgpool_fwd", r.y.ptr, r.out.ptr, N, st.res_in, st.res_in, r.y.C, st.down, r.y.ns,
                          r.out.ns, st.res_out)
            else:
                r.out = r.y
            e.blocks.append(r)
            e.acts[st.res_out] = r.out
            e.owner[st.res_out] = i
            cur = r.out
        return e

    def _decoder_fwd(self, prog: Program, N, pa: Dict[int, View], pa_sto: Dict[int, View], acts: Optional[Dict[int, View]],
                     given: Optional[Sequence[bool]] = None, want_z: bool = False, want_stats: bool = False,
                     explicit_eps: bool = True, kl_rows: Optional[torch.Tensor] = None,
                     kl_ch: Optional[torch.Tensor] = None,
                     z_views: Optional[Sequence[View]] = None, want_kl_elem: bool = False) -> Rec:
        """reference Decoder.forward (src/vae.py:222-301).  `given[i]` marks stochastic block i whose latent is
        supplied by the caller (forward_latents); other stochastic blocks sample q (acts given) or p.
        `z_views[i]`: the given latent already lives on the device as a bf16 planar view (another decoder pass of
        the same program produced it): used in place, no fp32 NCHW round trip."""
        dec = self.model.decoder
        zd = self.zd
        D = Rec()
        D.blocks = []
        D.latent_args = []
        D.eps = []
        D.z_in = {}
        D.z_out = {}
        D.stats_out = {}
        D.kl_elem = {}
        bias_of = {r: p for (r, _), p in zip(dec.bias_res, dec.bias)}
        w1 = dec.plan[0].cin
        h = new_act(N, 1, 1, w1, self.device)
        prog.call("cg_fill_planar", bias_of[1].data_ptr(), h.ptr, N, 1, w1, h.ns)
        zs = h  # h = z = bias[1].repeat (src/vae.py:232)
        cur_res = 1
        ksto = 0
        zfp_pending = False  # z_feat_proj of the previous block still running on the auxiliary lane
        for d in self.dec_layers:
            st = d.st
            r = Rec()
            r.d, r.st = d, st
            res = st.res
            r.up = None
            if zfp_pending:  # its output (the z stream) is consumed from here on
                prog.join()
                zfp_pending = False
            if cur_res < res:  # src/vae.py:251-262 (the z stream reuses the same bias b)
                b = bias_of.get(res)
                bptr = b.data_ptr() if b is not None else None
                r.up = (cur_res, h, zs, b)
                h_up = new_act(N, res, res, st.cin, self.device)
                prog.call("cg_upsample_fwd", h.ptr, bptr, h_up.ptr, N, cur_res, res, st.cin, h.ns, h_up.ns)
                if not self.q_corr:
                    if zs is h:
                        zs_up = h_up
                    else:
                        zs_up = new_act(N, res, res, st.cin, self.device)
                        prog.call("cg_upsample_fwd", zs.ptr, bptr, zs_up.ptr, N, cur_res, res, st.cin, zs.ns, zs_up.ns)
                    zs = zs_up
                h = h_up
                cur_res = res
            r.h_in, r.zs_in = h, zs
            # ---- posterior net (src/vae.py:185-192): shares only READ operands with the prior net, so it is emitted
            # first on the auxiliary lane and runs while the prior net occupies the main stream; they meet at the
            # latent kernel
            r.post, r.qstat = None, None
            if st.stochastic and acts is not None:
                r.qstat = View(torch.zeros(N, res, res, 2 * zd, device=self.device, dtype=torch.float32), 2 * zd)
                prog.fork()
                with prog.on_lane(1):
                    r.post = self._block_fwd(prog, d.post, [h, pa[res], acts[res]], N, res, res,
                                             final_segs=[SegSpec(r.qstat, 0)])
            # ---- prior (src/vae.py:172-183)
            p_src = [h if self.q_corr else zs] + ([pa_sto[res]] if self.cond_prior else [])
            # a block whose latent is GIVEN (forward_latents / the two decodes of a counterfactual) never reads p_loc /
            # p_logscale (src/vae.py:274-279): the prior conv then stores its feature columns only -- the fp32
            # statistics rows would be as many bytes again as the features at 96^2
            dead_stats = (st.stochastic and acts is None and given is not None and ksto < len(given)
                          and bool(given[ksto]) and not want_stats)
            r.pfeat = new_act(N, res, res, st.cin, self.device)
            if dead_stats:
                r.pstat = View(torch.zeros(1, 1, 1, 2 * zd, device=self.device, dtype=torch.float32), 2 * zd)  # placeholder
                segs = [SegSpec(r.pfeat, 2 * zd)]
            else:
                r.pstat = View(torch.zeros(N, res, res, 2 * zd, device=self.device, dtype=torch.float32), 2 * zd)
                segs = [SegSpec(r.pstat, 0), SegSpec(r.pfeat, 2 * zd)]
            r.prior = self._block_fwd(prog, d.prior, p_src, N, res, res, final_segs=segs)
            # ---- posterior + latent (src/vae.py:265-291)
            shared_z = (z_views is not None and st.stochastic and given is not None and ksto < len(given)
                        and bool(given[ksto]))
            r.z = z_views[ksto] if shared_z else new_act(N, res, res, zd, self.device)
            r.mode = 2
            r.eps = None
            la = None
            if st.stochastic:
                is_given = bool(given[ksto]) if given is not None and ksto < len(given) else False
                if acts is not None:
                    prog.join()  # posterior statistics ready
                    r.mode = 0
                elif is_given:
                    if not shared_z:
                        zin = torch.zeros(N, zd, res, res, device=self.device, dtype=torch.float32)
                        D.z_in[ksto] = zin
                        prog.call("cg_nchw_f32_to_planar", zin.data_ptr(), r.z.ptr, N, zd, res * res, r.z.ns)
                    r.mode = 3
                else:
                    r.mode = 1
                if r.mode in (0, 1):
                    la = L.LatentArgs()
                    la.p, la.p_ld = r.pstat.ptr, r.pstat.ns
                    if r.mode == 0:
                        la.q, la.q_ld = r.qstat.ptr, r.qstat.ns
                    if explicit_eps:
                        r.eps = torch.zeros(N, zd, res, res, device=self.device, dtype=torch.float32)
                        la.eps = r.eps.data_ptr()
                    D.eps.append(r.eps)
                    la.offset = ksto << 40
                    la.z_bf16, la.z_ns = r.z.ptr, r.z.ns
                    if want_z:
                        zo = torch.zeros(N, zd, res, res, device=self.device, dtype=torch.float32)
                        D.z_out[ksto] = zo
                        la.z_f32 = zo.data_ptr()
                    if kl_rows is not None and r.mode == 0:
                        la.kl_out = kl_rows[ksto].data_ptr()
                        if kl_ch is not None:  # kl_free_bits statistics (src/vae.py:443-449)
                            la.kl_ch = kl_ch[ksto].data_ptr()
                    if want_kl_elem and r.mode == 0:  # stats[i]["kl"] of the stand-alone Decoder call (src/vae.py:268)
                        ke = torch.zeros(N, zd, res, res, device=self.device, dtype=torch.float32)
                        D.kl_elem[ksto] = ke
                        la.kl_elem = ke.data_ptr()
                    la.N, la.HW, la.zdim, la.mode = N, res * res, zd, r.mode
                    if want_stats:
                        D.stats_out[ksto] = (r.qstat, r.pstat)
                ksto += 1
            else:
                la = L.LatentArgs()
                la.p, la.p_ld = r.pstat.ptr, r.pstat.ns
                la.z_bf16, la.z_ns = r.z.ptr, r.z.ns
                la.N, la.HW, la.zdim, la.mode = N, res * res, zd, 2
            if la is not None:
                prog.add(L.Launch("cg_latent_fwd", C.byref(la))).keep = (la, r)
                D.latent_args.append(la)
            # z stream of the next block (src/vae.py:297-300) only needs z and p_feat: auxiliary lane, joined at the
            # top of the next block
            r.zs_out = None
            if d.zfp is not None and st.idx + 1 < len(self.dec_layers):
                r.zs_out = new_act(N, res, res, st.cout, self.device)
                prog.fork()
                with prog.on_lane(1):
                    prog.add(d.zfp.forward([r.z, r.pfeat], [SegSpec(r.zs_out, 0)], N, res, res))
                zfp_pending = True
            # ---- merge (src/vae.py:292-300)
            r.h3 = new_act(N, res, res, st.cin, self.device)
            # h3 = h + p_feat + z_proj(cat[z, pa]) in one epilogue (src/vae.py:292-294)
            prog.add(d.z_proj.forward([r.z, pa[res]], [SegSpec(r.h3, 0, add=h, add2=r.pfeat)], N, res, res))
            r.conv = self._block_fwd(prog, d.conv, [r.h3], N, res, res)
            h = r.conv.y
            if r.zs_out is not None:
                zs = r.zs_out
            D.blocks.append(r)
        if zfp_pending:
            prog.join()
        D.h = h
        D.nsto = ksto
        return D

    # ================================================================== backward emission
    def _block_bwd(self, prog: Program, r: Rec, dy: View, dsrc: List[Optional[SegSpec]]):
        """Backward of _block_fwd.  dy: gradient wrt the last conv's full output (padded channels).
        dsrc[i]: where/how the first conv's data gradient for source i lands (None = not needed);
        mul/mul_act are filled in here.  Weight/bias gradients accumulate into the flat bucket."""
        bl = r.bl
        N, H, W = r.N, r.H, r.W
        ins = [r.srcs] + [[m] for m in r.mids_in]
        cur_dy = dy
        for li in range(len(bl.layers) - 1, -1, -1):
            layer = bl.layers[li]
            srcs = ins[li]
            prog.add(layer.wgrad(srcs, cur_dy, self.g(layer.weight), self.g(layer.bias), N, H, W))
            if li > 0:
                dmid = new_act(N, H, W, layer.src_logical[0], self.device)
                prog.add(layer.dgrad(0, cur_dy, SegSpec(dmid, 0, mul=r.mids[li - 1], mul_act=bl.act), N, H, W))
                cur_dy = dmid
            else:
                for i, sg in enumerate(dsrc):
                    if sg is None:
                        continue
                    sg.mul, sg.mul_act = srcs[i], bl.act
                    prog.add(layer.dgrad(i, cur_dy, sg, N, H, W))

    def _res_block_bwd(self, prog: Program, r: Rec, dout: View, extra: Optional[View] = None,
                       dx_out: Optional[View] = None) -> View:
        """Backward of a residual Block (encoder block / decoder `conv`), incl. pooling.
        Returns d(input) (written into `dx_out` when given).  `extra` is one more gradient contribution to the block
        input."""
        bl = r.bl
        N, H, W = r.N, r.H, r.W
        st = getattr(r, "st", None)
        dy = dout
        if st is not None and getattr(st, "down", None):
            dy = new_act(N, H, W, r.y.logical, self.device)
            prog.call("cg_avgpool_bwd", dout.ptr, dy.ptr, N, H, W, dy.C, st.down, dout.ns, dy.ns, st.res_out, 0)
        x = r.srcs[0]
        dx = dx_out if dx_out is not None else new_act(N, H, W, x.logical, self.device)
        if bl.proj is not None:
            skip = new_act(N, H, W, x.logical, self.device)
            prog.add(bl.proj.wgrad(r.srcs, dy, self.g(bl.proj.weight), self.g(bl.proj.bias), N, H, W))
            prog.add(bl.proj.dgrad(0, dy, SegSpec(skip, 0, add=extra), N, H, W))
        elif extra is not None:
            skip = new_act(N, H, W, x.logical, self.device)
            prog.call("cg_add", dy.ptr, extra.ptr, skip.ptr, N, H * W, skip.C, dy.ns, extra.ns, skip.ns)
        else:
            skip = dy
        self._block_bwd(prog, r, dy, [SegSpec(dx, 0, add=skip)])
        return dx

    def _decoder_bwd(self, prog: Program, D: Rec, dh_final: View, N, g_kl: float, acts_grad: Dict[int, View],
                     explicit_eps: bool, dz_extra: Optional[Dict[int, List[View]]] = None,
                     dz_out: Optional[Dict[int, View]] = None, kl_gate: Optional[torch.Tensor] = None):
        """Backward of _decoder_fwd.  `dz_extra[k]`: further gradients wrt the latent of stochastic block k (other decoder
        passes of the same program consumed it: counterfactual training).  `dz_out`: for passes whose latents were GIVEN
        (mode 3) the gradient wrt latent k is handed back here instead of going through a latent kernel."""
        dec = self.model.decoder
        zd = self.zd
        bias_param = {r: p for (r, _), p in zip(dec.bias_res, dec.bias)}
        dh_out, dzs_out = dh_final, None
        D.latent_bwd_args = []
        for r in reversed(D.blocks):
            d, st = r.d, r.st
            res = st.res
            # gradient of the prior's last conv output: [dp stats (2*zd) | d p_feat (cin)]
            DP = new_act(N, res, res, 2 * zd + st.cin, self.device)
            dz = new_act(N, res, res, zd, self.device)
            dz_written = False
            dpf = DP.slice(2 * zd, DP.C - 2 * zd, st.cin)
            # conv block.  Without a z_feat_proj consumer (last block) d p_feat IS d h3 (h3 = h + p_feat + ...): the block
            # input gradient is written straight into that slice instead of being copied there
            dh3 = self._res_block_bwd(prog, r.conv, dh_out, dx_out=dpf if r.zs_out is None else None)
            if r.zs_out is not None:
                prog.add(d.zfp.wgrad([r.z, r.pfeat], dzs_out, self.g(d.zfp.weight), self.g(d.zfp.bias), N, res, res))
                prog.add(d.zfp.dgrad(0, dzs_out, SegSpec(dz, 0), N, res, res))
                dz_written = True
                prog.add(d.zfp.dgrad(1, dzs_out, SegSpec(dpf, 0, add=dh3), N, res, res))
            # z_proj (no activation on its input)
            prog.add(d.z_proj.wgrad([r.z, r.pa], dh3, self.g(d.z_proj.weight), self.g(d.z_proj.bias), N, res, res))
            prog.add(d.z_proj.dgrad(0, dh3, SegSpec(dz, 0, add=dz if dz_written else None), N, res, res))
            if dz_extra is not None and st.stochastic:
                for ex in dz_extra.get(r.ksto, []):
                    prog.call("cg_add", dz.ptr, ex.ptr, dz.ptr, N, res * res, dz.C, dz.ns, ex.ns, dz.ns)
            if r.mode == 3:  # given latent: its gradient leaves the pass; the prior statistics were dead (DP stays 0 there)
                if dz_out is not None:
                    dz_out[r.ksto] = dz
                prog.keep.append((dz, DP))
            # latent
            lb = L.LatentBwdArgs()
            lb.p, lb.p_ld = r.pstat.ptr, r.pstat.ns
            dq = None
            if r.mode == 0:
                dq = new_act(N, res, res, 2 * zd, self.device)
                lb.q, lb.q_ld = r.qstat.ptr, r.qstat.ns
                lb.dq, lb.dq_ns = dq.ptr, dq.ns
                if explicit_eps:
                    lb.eps = r.eps.data_ptr()
                lb.offset = r.ksto << 40
            lb.dz, lb.dz_ns, lb.g_kl = dz.ptr, dz.ns, g_kl
            lb.dp, lb.dp_ns = DP.ptr, DP.ns
            lb.N, lb.HW, lb.zdim, lb.mode = N, res * res, zd, r.mode
            if kl_gate is not None and r.mode == 0:
                lb.kl_gate = kl_gate[r.ksto].data_ptr()
            if r.mode != 3:
                prog.add(L.Launch("cg_latent_bwd", C.byref(lb))).keep = (lb, dz, DP, dq)
                D.latent_bwd_args.append(lb)
            # posterior: d h_in = dh3 (+ posterior path), d acts[res] accumulates over blocks
            if r.post is not None:
                dh_in = new_act(N, res, res, st.cin, self.device)
                if res in acts_grad:
                    da = acts_grad[res]
                    seg_a = SegSpec(da, 0, add=da)
                else:
                    da = acts_grad[res] = new_act(N, res, res, st.cin, self.device)
                    seg_a = SegSpec(da, 0)
                # posterior and prior backward chains only meet again at the block input: the posterior chain runs on
                # the auxiliary lane (with q_correction the prior chain accumulates into the same tensor: no fork)
                forked = not self.q_corr
                if forked:
                    prog.fork()
                with prog.on_lane(1 if forked else 0):
                    self._block_bwd(prog, r.post, dq, [SegSpec(dh_in, 0, add=dh3), None, seg_a])
            else:
                forked = False
                dh_in = dh3
            # prior
            if self.q_corr:
                if dh_in is dh3:  # keep dh3 intact for clarity: accumulate into a fresh buffer
                    tmp = new_act(N, res, res, st.cin, self.device)
                    self._block_bwd(prog, r.prior, DP, [SegSpec(tmp, 0, add=dh_in)] + [None] * (len(r.prior.srcs) - 1))
                    dh_in = tmp
                else:
                    self._block_bwd(prog, r.prior, DP, [SegSpec(dh_in, 0, add=dh_in)] + [None] * (len(r.prior.srcs) - 1))
                dzs_in = None
            else:
                dzs_in = new_act(N, res, res, st.cin, self.device)
                self._block_bwd(prog, r.prior, DP, [SegSpec(dzs_in, 0)] + [None] * (len(r.prior.srcs) - 1))
            if forked:
                prog.join()
            # upsample
            if r.up is not None:
                src_res, h_prev, zs_prev, b = r.up
                dbias = self.g(b).data_ptr() if b is not None else None
                dh_prev = new_act(N, src_res, src_res, st.cin, self.device)
                prog.call("cg_upsample_bwd", dh_in.ptr, dh_prev.ptr, dbias, N, src_res, res, st.cin, dh_in.ns,
                          dh_prev.ns, 0)
                dzs_prev = None
                if dzs_in is not None:
                    if zs_prev is h_prev:  # first block: h and z are the same tensor
                        prog.call("cg_upsample_bwd", dzs_in.ptr, dh_prev.ptr, dbias, N, src_res, res, st.cin,
                                  dzs_in.ns, dh_prev.ns, 1)
                    else:
                        dzs_prev = new_act(N, src_res, src_res, st.cin, self.device)
                        prog.call("cg_upsample_bwd", dzs_in.ptr, dzs_prev.ptr, dbias, N, src_res, res, st.cin,
                                  dzs_in.ns, dzs_prev.ns, 0)
                dh_out, dzs_out = dh_prev, dzs_prev
            else:
                dh_out, dzs_out = dh_in, dzs_in
        # initial state h = z = bias[1] (src/vae.py:232)
        g1 = self.g(bias_param[1])
        w1 = dec.plan[0].cin
        prog.call("cg_colsum", dh_out.ptr, g1.data_ptr(), N, 1, w1, dh_out.ns)
        if dzs_out is not None:
            prog.call("cg_colsum", dzs_out.ptr, g1.data_ptr(), N, 1, w1, dzs_out.ns)

    def _encoder_bwd(self, prog: Program, e: Rec, acts_grad: Dict[int, View], N):
        enc = self.model.encoder
        din_next: Optional[View] = None
        for i in range(len(e.blocks) - 1, -1, -1):
            r = e.blocks[i]
            own = acts_grad.get(r.st.res_out) if e.owner.get(r.st.res_out) == i else None
            if din_next is None and own is None:
                continue  # nothing downstream of this block reaches the loss
            if din_next is None:
                dout, extra = own, None
            else:
                dout, extra = din_next, own
            if extra is not None:
                # d(out) = d(next block input) + d(acts[res]); fold the sum into one tensor first
                s = new_act(N, r.st.res_out, r.st.res_out, r.out.logical, self.device)
                prog.call("cg_add", dout.ptr, extra.ptr, s.ptr, N, r.st.res_out * r.st.res_out, s.C, dout.ns,
                          extra.ns, s.ns)
                dout = s
            din_next = self._res_block_bwd(prog, r, dout)
        if din_next is not None:
            prog.call("cg_stem_wgrad", e.x.data_ptr(), din_next.ptr, self.g(enc.stem.weight).data_ptr(),
                      self.g(enc.stem.bias).data_ptr(), N, self.C, self.R, self.args.widths[0], din_next.ns)

    # ================================================================== likelihood
    def _lik_args(self, h: View, x: Optional[torch.Tensor], N):
        lik = self.model.likelihood
        HW = self.R * self.R
        if self.dmol:
            a = L.DmolArgs()
            a.h, a.h_ns, a.Cw = h.ptr, h.ns, self.args.widths[0]
            a.x = x.data_ptr() if x is not None else None
            a.w, a.b = lik.conv.weight.data_ptr(), lik.conv.bias.data_ptr()
            a.N, a.HW = N, HW
            return a
        a = L.DGaussArgs()
        a.h, a.h_ns, a.Cw = h.ptr, h.ns, self.args.widths[0]
        a.x = x.data_ptr() if x is not None else None
        a.w_loc, a.b_loc = lik.x_loc.weight.data_ptr(), lik.x_loc.bias.data_ptr()
        a.w_ls, a.b_ls = lik.x_logscale.weight.data_ptr(), lik.x_logscale.bias.data_ptr()
        if self.C == 3:
            a.w_co, a.b_co = lik.channel_coeffs.weight.data_ptr(), lik.channel_coeffs.bias.data_ptr()
        a.N, a.HW, a.C = N, HW, self.C
        return a

    # ================================================================== programs
    def _inputs(self, prog: Program, N, with_x=True, n_pa=1):
        """static input buffers.  Parents arrive as (N, ctx); they are materialised per decoder resolution as
        spatially-constant planar tensors (= parents[..., :res, :res] of src/vae.py:241, in bf16)."""
        io = Rec()
        if with_x:
            io.x = torch.zeros(N, self.C, self.R, self.R, device=self.device, dtype=torch.float32)
        io.pa_in = [torch.zeros(N, self.ctx, device=self.device, dtype=torch.float32) for _ in range(n_pa)]
        io.pa, io.pa_sto = [], []
        # live hyper-parameters as DEVICE scalars {beta, p_sto scale of the conditioning dropout}: the kernels read them
        # at run time, so a captured CUDA graph follows beta annealing (src/trainer.py:52-57) and the per-step
        # dropout draw (src/vae.py:234-249) without re-capture
        io.hyp = torch.ones(4, device=self.device, dtype=torch.float32)
        io.hyp_host = torch.ones(4, dtype=torch.float32)  # pageable on purpose: the H2D copy snapshots it at call time
        resolutions = sorted({d.st.res for d in self.dec_layers})
        drop = self.model.decoder.is_drop_cond and self.cond_prior
        for t in io.pa_in:
            planes, planes_sto = {}, {}
            for res in resolutions:
                v = new_act(N, res, res, self.ctx, self.device)
                prog.call("cg_parents_plane", t.data_ptr(), self.ctx, 1, v.ptr, N, self.ctx, v.C, res * res, v.ns,
                          self.ctx, 1.0, None)
                planes[res] = v
                if drop:  # src/vae.py:244-247: channels 2: scaled by p_sto on the stochastic (prior) path
                    vs = new_act(N, res, res, self.ctx, self.device)
                    prog.call("cg_parents_plane", t.data_ptr(), self.ctx, 1, vs.ptr, N, self.ctx, vs.C, res * res,
                              vs.ns, 2, 1.0, io.hyp.data_ptr() + 4)
                    planes_sto[res] = vs
                else:
                    planes_sto[res] = v
            io.pa.append(planes)
            io.pa_sto.append(planes_sto)
        return io

    def dmol_mode(self) -> int:
        """cg_dmol_predict mode of DmolNet.mask (src/dmol.py:164-190): soft 0 | hard 1 | 'top<k>' 10 + k"""
        mask = getattr(self.model.likelihood, "mask", "soft")
        if mask == "soft":
            return 0
        if mask == "hard":
            return 1
        if "top" in mask:
            k = int(mask[-1])
            if not 0 < k < 10:
                raise ValueError("invalid top_k")  # src/dmol.py:180
            return 10 + k
        raise NotImplementedError(f"DmolNet.mask = {mask!r}")

    def _dmol_mode_arg(self, prog: Program):
        """mutable int32 launch argument, refreshed from DmolNet.mask whenever the program is fetched (HVAE._program)"""
        m = C.c_int32(self.dmol_mode())
        if not hasattr(prog, "dmol_modes"):
            prog.dmol_modes = []
        prog.dmol_modes.append(m)
        return m

    def build_elbo(self, N: int, train: bool, explicit_eps: bool) -> Program:
        """HVAE.forward (src/vae.py:439-458) and, when `train`, its full backward"""
        prog = Program(f"elbo(N={N},train={train})")
        io = self._inputs(prog, N)
        prog.io = io
        nsto = sum(1 for d in self.dec_layers if d.st.stochastic)
        prog.kl_rows = torch.zeros(max(nsto, 1), N, device=self.device, dtype=torch.float32)
        prog.nll = torch.zeros(N, device=self.device, dtype=torch.float32)
        prog.out3 = torch.zeros(3, device=self.device, dtype=torch.float32)
        prog.zero = [prog.kl_rows, prog.nll]
        fb = float(self.model.free_bits)
        prog.kl_ch = prog.kl_gate = None
        if fb > 0:  # src/vae.py:443-449: the floor acts on per-channel batch means, so the forward also collects those
            prog.kl_ch = torch.zeros(max(nsto, 1), 16, device=self.device, dtype=torch.float32)
            prog.kl_gate = torch.zeros(max(nsto, 1), 16, device=self.device, dtype=torch.float32)
            prog.kl_fb_row = torch.zeros(N, device=self.device, dtype=torch.float32)
            prog.zero.append(prog.kl_ch)
        e = self._encoder_fwd(prog, io.x, N)
        D = self._decoder_fwd(prog, N, io.pa[0], io.pa_sto[0], e.acts, explicit_eps=explicit_eps,
                              kl_rows=prog.kl_rows, kl_ch=prog.kl_ch)
        k = 0
        for r in D.blocks:
            r.pa = io.pa[0][r.st.res]
            r.ksto = k
            if r.st.stochastic:
                k += 1
        prog.D, prog.e = D, e
        prog.eps = D.eps
        la = self._lik_args(D.h, io.x, N)
        la.nll = prog.nll.data_ptr()
        prog.lik = la
        prog.add(L.Launch("cg_dmol_loss_fwd" if self.dmol else "cg_dgauss_nll_fwd", C.byref(la)))
        npix = float(self.C * self.R * self.R)
        if fb > 0:
            # data parallel: the batch mean runs over the GLOBAL batch -- sum the 16 x nsto statistics over ranks before the
            # max; a host-side collective, so such a program is not captured into a CUDA graph
            from . import dp
            world, _ = dp.world_info()
            if world > 1:
                prog.add(HostOp("allreduce_kl_ch", lambda s, t=prog.kl_ch: dp.all_reduce_sum_(t)))
                prog.no_graph = True
            prog.call("cg_free_bits", prog.kl_ch.data_ptr(), max(nsto, 1), fb, 1.0 / (N * world),
                      prog.kl_gate.data_ptr(), prog.kl_fb_row.data_ptr(), N)
            prog.fin = prog.call("cg_elbo_finalize", prog.nll.data_ptr(), prog.kl_fb_row.data_ptr(), prog.out3.data_ptr(),
                                 N, 1, 1.0 / npix, 1.0, io.hyp.data_ptr())
        else:
            prog.fin = prog.call("cg_elbo_finalize", prog.nll.data_ptr(), prog.kl_rows.data_ptr(), prog.out3.data_ptr(),
                                 N, max(nsto, 1), 1.0 / npix, 1.0, io.hyp.data_ptr())
        prog.n_fwd = len(prog.launches)
        if train:
            prog.beta_users = []
            lik = self.model.likelihood
            dh = new_act(N, self.R, self.R, self.args.widths[0], self.device)
            lb = self._lik_args(D.h, io.x, N)
            lb.g = 1.0 / N
            lb.dh, lb.dh_ns = dh.ptr, dh.ns
            if self.dmol:
                lb.dw, lb.db = self.g(lik.conv.weight).data_ptr(), self.g(lik.conv.bias).data_ptr()
                prog.add(L.Launch("cg_dmol_loss_bwd", C.byref(lb))).keep = (lb, dh)
            else:
                lb.dw_loc, lb.db_loc = self.g(lik.x_loc.weight).data_ptr(), self.g(lik.x_loc.bias).data_ptr()
                lb.dw_ls, lb.db_ls = self.g(lik.x_logscale.weight).data_ptr(), self.g(lik.x_logscale.bias).data_ptr()
                if self.C == 3:
                    lb.dw_co = self.g(lik.channel_coeffs.weight).data_ptr()
                    lb.db_co = self.g(lik.channel_coeffs.bias).data_ptr()
                prog.add(L.Launch("cg_dgauss_nll_bwd", C.byref(lb))).keep = (lb, dh)
            acts_grad: Dict[int, View] = {}
            # free bits: d kl / d KL[b,c,h,w] = gate[c] / (global batch) = gate[c] / N after the data-parallel 1/world that
            # the optimiser applies to the summed bucket -- the same per-rank coefficient as without free bits
            self._decoder_bwd(prog, D, dh, N, 1.0 / (N * npix), acts_grad, explicit_eps, kl_gate=prog.kl_gate)
            for lb in D.latent_bwd_args:  # g_kl = (1 / (N * npix)) * live beta
                lb.g_kl_dev = io.hyp.data_ptr()
            self._encoder_bwd(prog, e, acts_grad, N)
            prog.npix = npix
        return prog

    @staticmethod
    def set_hyper(prog: Program, beta: Optional[float] = None, drop_sto: Optional[float] = None):
        """update the program's device scalars (stream-ordered 16-byte copy, issued only when a value changed; never
        inside a graph capture -- captured programs read the scalars at replay time)"""
        io = prog.io
        new = io.hyp_host.clone()
        if beta is not None:
            new[0] = float(beta)
        if drop_sto is not None:
            new[1] = float(drop_sto)
        if not torch.equal(new, io.hyp_host):
            io.hyp_host.copy_(new)
            io.hyp.copy_(io.hyp_host)

    def build_decode(self, N: int, kind: str, given: Optional[Sequence[bool]] = None, n_pa: int = 1,
                     want_stats: bool = False) -> Program:
        """kind: 'abduct' (encoder + posterior pass returning z), 'latents' (forward_latents / sample)"""
        prog = Program(f"{kind}(N={N})")
        if kind == "abduct":
            io = self._inputs(prog, N)
            e = self._encoder_fwd(prog, io.x, N)
            prog.D = self._decoder_fwd(prog, N, io.pa[0], io.pa_sto[0], e.acts, want_z=True, want_stats=want_stats)
            prog.io = io
            return prog
        io = self._inputs(prog, N, with_x=False, n_pa=n_pa)
        prog.io = io
        prog.Ds, prog.x_out, prog.scale_out, prog.lik_args = [], [], [], []
        for j in range(n_pa):
            D = self._decoder_fwd(prog, N, io.pa[j], io.pa_sto[j], None, given=given, want_stats=want_stats)
            xo = torch.zeros(N, self.C, self.R, self.R, device=self.device, dtype=torch.float32)
            so = torch.zeros_like(xo)
            la = self._lik_args(D.h, None, N)
            if self.dmol:
                prog.add(L.Launch("cg_dmol_predict", C.byref(la), self._dmol_mode_arg(prog), None, None, C.c_float(0.0), xo.data_ptr(),
                                  so.data_ptr()))
            else:
                prog.add(L.Launch("cg_dgauss_sample", C.byref(la), xo.data_ptr(), so.data_ptr(), None, C.c_float(0.0)))
            prog.Ds.append(D)
            prog.x_out.append(xo)
            prog.scale_out.append(so)
            prog.lik_args.append(la)
        return prog

    # ---- stand-alone calls of the sub-modules (reference surface: model.encoder(x), model.decoder(...),
    # model.likelihood.nll / .sample).  fp32 NCHW at the boundary like the reference; inference only.
    def _nchw_out(self, prog: Program, v: View, N, C_, res) -> torch.Tensor:
        out = torch.zeros(N, C_, res, res, device=self.device, dtype=torch.float32)
        prog.call("cg_planar_to_nchw_f32", v.ptr, out.data_ptr(), N, C_, res * res, v.ns)
        return out

    def _nchw_in(self, prog: Program, N, C_, res):
        buf = torch.zeros(N, C_, res, res, device=self.device, dtype=torch.float32)
        v = new_act(N, res, res, C_, self.device)
        prog.call("cg_nchw_f32_to_planar", buf.data_ptr(), v.ptr, N, C_, res * res, v.ns)
        return buf, v

    def build_encoder_call(self, N: int) -> Program:
        """Encoder.forward (src/vae.py:124-134): {resolution: activations}"""
        prog = Program(f"encoder(N={N})")
        prog.x = torch.zeros(N, self.C, self.R, self.R, device=self.device, dtype=torch.float32)
        e = self._encoder_fwd(prog, prog.x, N)
        prog.acts_out = {res: self._nchw_out(prog, v, N, v.logical, res) for res, v in e.acts.items()}
        prog.e = e
        return prog

    def build_decoder_call(self, N: int, has_acts: bool, given: Optional[Sequence[bool]]) -> Program:
        """Decoder.forward (src/vae.py:222-301) on caller-supplied encoder activations / latents"""
        prog = Program(f"decoder(N={N},acts={has_acts})")
        io = self._inputs(prog, N, with_x=False, n_pa=1)
        prog.io = io
        prog.acts_in, acts = {}, None
        if has_acts:
            acts = {}
            for res, st in {st.res_out: st for st in self.model.encoder.plan}.items():  # last block of each resolution
                prog.acts_in[res], acts[res] = self._nchw_in(prog, N, st.cout, res)
        prog.D = self._decoder_fwd(prog, N, io.pa[0], io.pa_sto[0], acts, given=given, want_z=has_acts, want_stats=True,
                                   want_kl_elem=has_acts)
        prog.h_out = self._nchw_out(prog, prog.D.h, N, self.args.widths[0], self.R)
        return prog

    def build_likelihood_call(self, N: int, kind: str) -> Program:
        """DGaussNet / DmolNet .nll(h, x) and .sample(h) (src/vae.py:352-422, src/dmol.py:228-245) on caller-supplied h"""
        prog = Program(f"likelihood.{kind}(N={N})")
        prog.h_in, hv = self._nchw_in(prog, N, self.args.widths[0], self.R)
        if kind == "nll":
            prog.x = torch.zeros(N, self.C, self.R, self.R, device=self.device, dtype=torch.float32)
            prog.nll = torch.zeros(N, device=self.device, dtype=torch.float32)
            la = self._lik_args(hv, prog.x, N)
            la.nll = prog.nll.data_ptr()
            prog.add(L.Launch("cg_dmol_loss_fwd" if self.dmol else "cg_dgauss_nll_fwd", C.byref(la))).keep = (la, hv)
            prog.zero = [prog.nll]
        else:
            prog.x_out = torch.zeros(N, self.C, self.R, self.R, device=self.device, dtype=torch.float32)
            prog.scale_out = torch.zeros_like(prog.x_out)
            la = self._lik_args(hv, None, N)
            if self.dmol:
                prog.add(L.Launch("cg_dmol_predict", C.byref(la), self._dmol_mode_arg(prog), None, None, C.c_float(0.0),
                                  prog.x_out.data_ptr(), prog.scale_out.data_ptr())).keep = (la, hv)
            else:
                prog.add(L.Launch("cg_dgauss_sample", C.byref(la), prog.x_out.data_ptr(), prog.scale_out.data_ptr(), None,
                                  C.c_float(0.0))).keep = (la, hv)
        return prog

    def build_counterfactual(self, N: int, train: bool = False) -> Program:
        """DSCM.forward hot lines (src/pgm/dscm.py:52-56) as ONE program: encoder + posterior decoder pass (abduction,
        in-kernel Philox noise), two prior-only decoder passes on the abducted latents (counterfactual and observed
        parents) reading the bf16 latents of the first pass in place, likelihood means, and the combine kernel.
        `train`: also records `prog.bwd`, the hand-derived backward of the whole pass given d cf_x (`prog.dcf`): the
        reference back-propagates aux_loss(cf_x) into the HVAE (src/pgm/dscm.py:78-88, src/pgm/train_cf.py:159-180)."""
        prog = Program(f"counterfactual(N={N},train={train})")
        io = self._inputs(prog, N, with_x=True, n_pa=2)  # pa_in[0]: observed parents, pa_in[1]: counterfactual parents
        prog.io = io
        e = self._encoder_fwd(prog, io.x, N)
        Da = self._decoder_fwd(prog, N, io.pa[0], io.pa_sto[0], e.acts, explicit_eps=False)
        prog.D = Da
        prog.seed_ctr = torch.zeros(1, dtype=torch.int64, device=self.device)
        for la in Da.latent_args:
            la.seed_dev = prog.seed_ctr.data_ptr()
        zs = [r.z for r in Da.blocks if r.st.stochastic]
        given = [True] * len(zs)
        outs, passes = [], []
        for j in (1, 0):
            D = self._decoder_fwd(prog, N, io.pa[j], io.pa_sto[j], None, given=given, z_views=zs)
            xo = torch.zeros(N, self.C, self.R, self.R, device=self.device, dtype=torch.float32)
            so = torch.zeros_like(xo)
            la = self._lik_args(D.h, None, N)
            if self.dmol:
                prog.add(L.Launch("cg_dmol_predict", C.byref(la), self._dmol_mode_arg(prog), None, None, C.c_float(0.0), xo.data_ptr(),
                                  so.data_ptr())).keep = (la, D)
            else:
                prog.add(L.Launch("cg_dgauss_sample", C.byref(la), xo.data_ptr(), so.data_ptr(), None,
                                  C.c_float(0.0))).keep = (la, D)
            outs.append((xo, so))
            passes.append((D, j))
        (cf_loc, cf_scale), (rec_loc, rec_scale) = outs
        prog.cf_x = torch.zeros_like(cf_loc)
        prog.acc = torch.zeros_like(cf_loc)
        prog.acc2 = torch.zeros_like(cf_loc)
        prog.call("cg_cf_combine", io.x.data_ptr(), rec_loc.data_ptr(), rec_scale.data_ptr(), cf_loc.data_ptr(),
                  cf_scale.data_ptr(), prog.cf_x.data_ptr(), prog.acc.data_ptr(), prog.acc2.data_ptr(), io.x.numel())
        prog.keep += [outs]
        if not train:
            return prog
        if self.dmol:
            raise NotImplementedError("gradients through the counterfactual are implemented for the DGaussNet likelihood "
                                      "(every reference counterfactual-training config, src/pgm/train_cf.py); DmolNet: no")
        # ------------------------------------------------------------------ backward program (given prog.dcf = d loss / d cf_x)
        bwd = Program(f"counterfactual_bwd(N={N})")
        prog.bwd = bwd
        prog.dcf = torch.zeros_like(cf_loc)
        dl = [torch.zeros_like(cf_loc) for _ in range(4)]  # d rec_loc, d rec_scale, d cf_loc, d cf_scale
        bwd.call("cg_cf_combine_bwd", io.x.data_ptr(), rec_loc.data_ptr(), rec_scale.data_ptr(), cf_loc.data_ptr(),
                 cf_scale.data_ptr(), prog.dcf.data_ptr(), dl[0].data_ptr(), dl[1].data_ptr(), dl[2].data_ptr(),
                 dl[3].data_ptr(), io.x.numel())
        bwd.keep += [dl]
        lik = self.model.likelihood
        dz_extra: Dict[int, List[View]] = {}
        for (D, j), (dloc, dscale) in zip(passes, ((dl[2], dl[3]), (dl[0], dl[1]))):
            k = 0
            for r in D.blocks:
                r.pa, r.ksto = io.pa[j][r.st.res], k
                if r.st.stochastic:
                    k += 1
            dh = new_act(N, self.R, self.R, self.args.widths[0], self.device)
            lb = self._lik_args(D.h, None, N)
            lb.dh, lb.dh_ns = dh.ptr, dh.ns
            lb.dw_loc, lb.db_loc = self.g(lik.x_loc.weight).data_ptr(), self.g(lik.x_loc.bias).data_ptr()
            lb.dw_ls, lb.db_ls = self.g(lik.x_logscale.weight).data_ptr(), self.g(lik.x_logscale.bias).data_ptr()
            if self.C == 3:
                lb.dw_co = self.g(lik.channel_coeffs.weight).data_ptr()
                lb.db_co = self.g(lik.channel_coeffs.bias).data_ptr()
            bwd.add(L.Launch("cg_dgauss_sample_bwd", C.byref(lb), dloc.data_ptr(), dscale.data_ptr())).keep = (lb, dh)
            dz_pass: Dict[int, View] = {}
            self._decoder_bwd(bwd, D, dh, N, 0.0, {}, False, dz_out=dz_pass)
            for ks, v in dz_pass.items():
                dz_extra.setdefault(ks, []).append(v)
        # abduction pass: its output h feeds nothing (zero upstream gradient); the latents carry everything
        k = 0
        for r in Da.blocks:
            r.pa, r.ksto = io.pa[0][r.st.res], k
            if r.st.stochastic:
                k += 1
        dh0 = new_act(N, self.R, self.R, self.args.widths[0], self.device)  # stays zero
        acts_grad: Dict[int, View] = {}
        self._decoder_bwd(bwd, Da, dh0, N, 0.0, acts_grad, False, dz_extra=dz_extra)
        for lb in Da.latent_bwd_args:
            lb.seed_dev = prog.seed_ctr.data_ptr()
        self._encoder_bwd(bwd, e, acts_grad, N)
        return prog
