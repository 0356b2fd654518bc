"""Regenerates the algorithmic-work table of SURVEY.md section 8(d) from the presets and arch strings alone
(no torch hooks): conv FLOPs = sum 2*Cin*Cout*k^2*Ho*Wo over every nn.Conv2d executed, "layerwise bytes" = sum of
bf16 (input + output) activation bytes per conv.  Usage: python tools/roofline.py [config ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "causal-gen_b200"))
from causalgen_b200.arch import decoder_plan, encoder_plan
from causalgen_b200.presets import PRESETS, make_args


def block_convs(cin, cmid, cout, k, light, residual, proj):
    """(cin, cout, k) of every conv of a reference Block (src/vae.py:49-78)"""
    if light:
        convs = [(cin, cmid, k), (cmid, cout, k)]
    else:
        convs = [(cin, cmid, 1), (cmid, cmid, k), (cmid, cmid, k), (cmid, cout, 1)]
    if residual and proj:
        convs.append((cin, cout, 1))
    return convs


def count(args, passes):
    """passes: subset of {'enc', 'post', 'prior'}; returns (flops, bytes) per image"""
    light = "ukbb" in args.hps
    zd, ctx = args.z_dim, args.context_dim
    fl = by = 0

    def add(convs, res):
        nonlocal fl, by
        for ci, co, k in convs:
            fl += 2 * ci * co * k * k * res * res
            by += 2 * (ci + co) * res * res

    if "enc" in passes:
        R = args.input_res
        add([(args.input_channels, args.widths[0], 7)], R)
        for st in encoder_plan(args):
            add(block_convs(st.cin, st.cmid, st.cout, 3, light, True, bool(st.down) or st.cin > st.cout), st.res_in)
    plan = decoder_plan(args)
    for st in plan:
        k = st.ksize
        pin = st.cin + (ctx if args.cond_prior else 0)
        add(block_convs(pin, st.cmid, 2 * zd + st.cin, k, light, False, False), st.res)          # prior
        if st.stochastic and "post" in passes:
            add(block_convs(2 * st.cin + ctx, st.cmid, 2 * zd, k, light, False, False), st.res)  # posterior
        add([(zd + ctx, st.cin, 1)], st.res)                                                     # z_proj
        if not args.q_correction and st.idx + 1 < len(plan):
            add([(zd + st.cin, st.cout, 1)], st.res)                                             # z_feat_proj
        add(block_convs(st.cin, st.cmid, st.cout, k, light, True, st.cin > st.cout), st.res)     # conv
    add([(args.widths[0], args.input_channels, 1)] * (2 if args.input_channels == 1 else 3), args.input_res)
    return fl, by


if __name__ == "__main__":
    names = sys.argv[1:] or ["morphomnist", "cmnist", "ukbb192", "mimic192", "mimic224"]
    print("%-12s %12s %14s %12s %12s %12s %14s" % ("config", "fwd GFLOP", "train GFLOP", "abduct", "fwd_latents", "CF GFLOP",
                                                    "fwd bytes MB"))
    for n in names:
        a = make_args(n)
        f_fwd, b_fwd = count(a, {"enc", "post", "prior"})
        f_lat, _ = count(a, {"prior"})
        print("%-12s %12.4f %14.4f %12.4f %12.4f %12.4f %14.1f" % (n, f_fwd / 1e9, 3 * f_fwd / 1e9, f_fwd / 1e9, f_lat / 1e9,
                                                                  (f_fwd + 2 * f_lat) / 1e9, b_fwd / 1e6))
