"""Diagnostic: every intermediate of one TCN block forward/backward vs the fp64 oracle."""
import os, sys, json
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from oracle import spexplus as ospex
from wesep_b200 import ops, synth
from wesep_b200.modules.tasnet.convs import Conv1DBlock

DEV = "cuda"


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(DEV)


def run(n, B, H, T, dil, seed):
    blk = Conv1DBlock(B, H, 3, dil, "gLN", False, False)
    synth.fill_state_dict_(blk.state_dict(), seed=seed)
    blk = blk.to(DEV)
    x = ops.new_act(n, B, T, DEV); x.copy_(rnd(n, B, T, seed=seed + 1)); x.requires_grad_(True)
    gy = rnd(n, B, T, seed=seed + 3)
    res = {}
    for rep in range(2):
        ops.DEBUG_STASH = {}
        out = blk(x)
        grads = torch.autograd.grad(out, [x] + list(blk.parameters()), gy)
        st = ops.DEBUG_STASH
        if rep == 0:
            first = [g.clone() for g in grads]
            first_dd = st["dd"].clone()
        else:
            res["nondeterminism_dx"] = rel(grads[0], first[0])
            res["nondeterminism_dd"] = rel(st["dd"], first_dd)
    sd = {k: v.detach().double().requires_grad_(True) for k, v in blk.state_dict().items()}
    x64 = x.detach().double().requires_grad_(True)
    u = F.conv1d(x64, sd["conv1x1.weight"], sd["conv1x1.bias"]); u.retain_grad()
    y1 = F.prelu(u, sd["PReLU_1.weight"])
    z1 = ospex.gln(y1, sd["norm_1.weight"], sd["norm_1.bias"]); z1.retain_grad()
    d = F.conv1d(z1, sd["dwconv.weight"], sd["dwconv.bias"], padding=dil, dilation=dil, groups=H); d.retain_grad()
    y2 = F.prelu(d, sd["PReLU_2.weight"])
    z2 = ospex.gln(y2, sd["norm_2.weight"], sd["norm_2.bias"])
    o = F.conv1d(z2, sd["Output.weight"], sd["Output.bias"]) + x64
    names = [k for k, _ in blk.named_parameters()]
    o.backward(gy.double())
    ref = [x64.grad] + [sd[k].grad for k in names]
    res["alpha"] = [float(sd["PReLU_1.weight"]), float(sd["PReLU_2.weight"])]
    res["out"] = rel(out, o); res["u"] = rel(st["u"], u); res["d"] = rel(st["d"], d)
    res["dd"] = rel(st["dd"], d.grad); res["du"] = rel(st["du"], u.grad)
    res["dd_rows"] = [rel(st["dd"][i], d.grad[i]) for i in range(n)]
    res["du_rows"] = [rel(st["du"][i], u.grad[i]) for i in range(n)]
    M = H * T
    g2 = sd["norm_2.weight"].detach().view(-1); W3 = sd["Output.weight"].detach().view(B, H)
    h = torch.einsum("oc,not->nct", W3, gy.double()) * g2[None, :, None]
    mu2 = y2.detach().mean((1, 2), keepdim=True); var2 = ((y2.detach() - mu2) ** 2).mean((1, 2), keepdim=True)
    yh2 = (y2.detach() - mu2) / torch.sqrt(var2 + 1e-5)
    res["mh_ref"] = h.mean((1, 2)).tolist(); res["mh"] = st["rowsc"][:, 0].tolist()
    res["mhy_ref"] = (h * yh2).mean((1, 2)).tolist(); res["mhy"] = st["rowsc"][:, 1].tolist()
    res["mu2_ref"] = mu2.flatten().tolist(); res["mu2"] = st["rowsc"][:, 6].tolist()
    res["r2_ref"] = (1 / torch.sqrt(var2 + 1e-5)).flatten().tolist(); res["r2"] = st["rowsc"][:, 7].tolist()
    g1 = sd["norm_1.weight"].detach().view(-1)
    dz1 = z1.grad
    mu1 = y1.detach().mean((1, 2), keepdim=True); var1 = ((y1.detach() - mu1) ** 2).mean((1, 2), keepdim=True)
    yh1 = (y1.detach() - mu1) / torch.sqrt(var1 + 1e-5)
    h1 = dz1 * g1[None, :, None]
    res["m1_ref"] = h1.mean((1, 2)).tolist(); res["m1"] = (st["rowsc"][:, 2] / M).tolist()
    res["m2_ref"] = (h1 * yh1).mean((1, 2)).tolist(); res["m2"] = ((st["rowsc"][:, 3] - st["rowsc"][:, 4]) / M).tolist()
    res["sg"] = rel(st["sg"], gy.double().sum(-1))
    res["Gn"] = rel(st["Gn"], torch.einsum("not,nct->noc", gy.double(), y2.detach()))
    for nm, a, b in zip(["dx"] + names, grads, ref):
        res["g/" + nm] = rel(a, b)
    e = (grads[0].double() - ref[0]).abs()
    res["dx_rows"] = [rel(grads[0][i], ref[0][i]) for i in range(n)]
    ed = (st["dd"].double() - d.grad).abs()
    res["dd_outliers"] = torch.nonzero(ed > 50 * ed.mean())[:12].tolist()
    res["dd_outlier_count"] = int((ed > 50 * ed.mean()).sum())
    res["dd_err_max_mean"] = [float(ed.max()), float(ed.mean()), float(d.grad.abs().mean())]
    sdu_ref = u.grad.sum(-1)
    res["sdu"] = rel(st["sdu"], sdu_ref)
    res["sdu_rows"] = [rel(st["sdu"][i], sdu_ref[i]) for i in range(n)]
    errdu = (st["du"].double() - u.grad)
    res["du_signed_mean_err_rows"] = errdu.mean((1, 2)).tolist()
    res["du_abs_mean_rows"] = u.grad.abs().mean((1, 2)).tolist()
    # error by t-tile of 128 and by channel (row 0)
    e0 = errdu[0]
    tt = [float(e0[:, i:i + 128].norm() / (u.grad[0][:, i:i + 128].norm() + 1e-30)) for i in range(0, T, max(128, T // 16 // 128 * 128))]
    res["du_row0_err_by_t"] = tt
    ce = (e0.norm(dim=1) / (u.grad[0].norm(dim=1) + 1e-30))
    res["du_row0_err_by_ch_top"] = sorted([(float(v), i) for i, v in enumerate(ce.tolist())], reverse=True)[:6]
    res["du_row0_err_by_ch_median"] = float(ce.median())
    # is the error explained by a scalar offset in m1 / m2?  du_err ~ -r1*(dm1 + yh1*dm2)*slope
    a1 = float(sd["PReLU_1.weight"])
    slope = torch.where(u.detach() > 0, torch.ones_like(u.detach()), torch.full_like(u.detach(), a1))
    r1 = (1 / torch.sqrt(var1 + 1e-5))
    A = torch.stack([(-(r1 * slope))[0].flatten(), (-(r1 * slope * yh1))[0].flatten()], 1)
    sol = torch.linalg.lstsq(A, e0.flatten()[:, None]).solution.flatten()
    res["fit_dm1_dm2_row0"] = sol.tolist()
    res["fit_residual_rel"] = float((A @ sol - e0.flatten()).norm() / (e0.norm() + 1e-30))
    edu = errdu.abs()
    res["du_outliers"] = torch.nonzero(edu > 50 * edu.mean())[:12].tolist()
    res["du_outlier_count"] = int((edu > 50 * edu.mean()).sum())
    return res


if __name__ == "__main__":
    out = {}
    for cfg in [(2, 64, 128, 1000, 2, 12), (2, 64, 128, 1000, 4, 14), (2, 256, 512, 6399, 128, 7), (1, 64, 128, 1000, 2, 12),
                (3, 64, 128, 1000, 4, 12)]:
        out[str(cfg)] = run(*cfg)
        print(cfg, json.dumps(out[str(cfg)]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "diag_tcn.json"), "w"), indent=1)
