"""A/B timing of one Spex+ TCN block (fwd + bwd, n rows, dil 8) under several library builds on the SAME box.

  python tools/ab_block.py libA.so libB.so ...     (paths relative to the repo root; '-' = the default build)

Each build runs in its own subprocess (the library is loaded once per process) and reports CUPTI kernel times
averaged over `reps` passes; the order is repeated twice (A B A B) so that clock / thermal drift shows up.
"""
import os, sys, subprocess, collections
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import torch
    from torch.profiler import profile, ProfilerActivity
    from wesep_b200 import ops, synth
    from wesep_b200.modules.tasnet.convs import Conv1DBlock
    n, dil, reps = 32, 8, 6
    if os.environ.get("AB_FLAGS"):
        from wesep_b200 import _lib
        _lib.lib().wesep_b200_set_tc_flags(int(os.environ["AB_FLAGS"]))
    blk = Conv1DBlock(256, 512, 3, dil, "gLN", False, False)
    synth.fill_state_dict_(blk.state_dict(), seed=1)
    blk = blk.to("cuda")
    x = ops.new_act(n, 256, 6399, "cuda"); x.normal_(); x.requires_grad_(True)
    g = ops.new_act(n, 256, 6399, "cuda"); g.normal_()

    def one():
        y = blk(x)
        torch.autograd.grad(y, [x] + list(blk.parameters()), g)
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        one()
    e.record(); torch.cuda.synchronize()
    total = s.elapsed_time(e) / reps * 1e3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            one()
        torch.cuda.synchronize()
    agg = collections.defaultdict(float)
    for ev in prof.events():
        if ev.device_type.name == "CUDA" and ev.name.startswith("void wb::") or ev.name.startswith("wb::"):
            nm = ev.name.split("(")[0].replace("void ", "").replace("wb::", "")
            agg[nm] += (ev.time_range.end - ev.time_range.start) / reps
    keys = sorted(agg, key=lambda k: -agg[k])[:9]
    print(f"block fwd+bwd {total:7.1f} us | " + "  ".join(f"{k[:26]} {agg[k]:.0f}" for k in keys), flush=True)
    sys.exit(0)

libs = sys.argv[1:] or ["-"]
for rnd in range(2):
    for lib in libs:
        env = dict(os.environ)
        if "@" in lib:                      # "path@flags": wesep_b200_set_tc_flags(flags) in the child
            lib_path, env["AB_FLAGS"] = lib.split("@")
        else:
            lib_path = lib
        if lib_path != "-":
            env["WESEP_B200_LIB"] = os.path.join(ROOT, lib_path)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True, timeout=300)
        out = [l for l in r.stdout.splitlines() if l.startswith("block")]
        print(f"[{lib}] " + (out[0] if out else "FAILED: " + r.stderr[-400:]), flush=True)
