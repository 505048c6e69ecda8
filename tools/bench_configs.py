"""Step times of the BASELINE.json configs on one B200 (ours, eager and CUDA-graph replay) next to (a) the CPU oracle port and (b) -- informational,
SURVEY 2.2's bar "beat aten / cuBLAS on the same box" -- the SAME oracle functions run eagerly on the GPU (torch ops = aten + cuBLAS + cuDNN).
cfg0 ABMIL gated fwd/bwd N=1024; cfg1 MHIM(attn) N=10k (teacher, student fwd, bwd); cfg2 MHIM(selfattn) N=50k forward_test /
teacher; cfg3 MHIM(dsmil) N=10k D=1536; cfg4 is tools/bench_sharded.py (multi-GPU)."""
import json, os, sys, time
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases, mhimk
from mhimk import modules as M
from oracle import mil_oracle as O

dev = torch.device("cuda")
LABEL = torch.tensor([1], device=dev)
REPS = int(os.environ.get("CFG_REPS", 10))
CPU = os.environ.get("CFG_CPU", "1") == "1"


def gpu_time(fn, reps=REPS):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def cpu_time(fn, reps=3):
    torch.set_num_threads(os.cpu_count())
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return sorted(ts)[len(ts) // 2] * 1e3


out = {}
# ---- cfg0: gated ABMIL fwd+bwd, N=1024
sd = cases.gated_state(1)
g = M.AttentionGated(1024, 2, act="relu", dropout=0.0).to(dev).train(); g.load_state_dict({k: v.to(dev) for k, v in sd.items()})
x = cases.make_bag(2, 1024, 1024).to(dev)
def step0():
    g.zero_grad(set_to_none=True)
    F.cross_entropy(g(x), LABEL).backward()
out["cfg0_gated_abmil_fwd_bwd_N1024_ms"] = gpu_time(step0)
if CPU:
    sdl = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xc = x.cpu()
    def cpu0():
        for v in sdl.values(): v.grad = None
        F.cross_entropy(O.abmil_gated(sdl, xc, "relu"), torch.tensor([1])).backward()
    out["cfg0_cpu_ms"] = cpu_time(cpu0)

# ---- cfg1: MHIM(attn) N=10k D=1024, teacher + student fwd + bwd
def mhim_cfg(base, N, D, tag, do_cpu=True, student=True):
    kw = dict(cases.MHIM_KW, baseline=base, input_dim=D, dropout=0.0)
    stu, tea = M.MHIM(**kw).to(dev).train(), M.MHIM(**kw).to(dev).train()
    stu.load_state_dict({k: v.to(dev) for k, v in cases.mhim_state(1, base, D=D).items()})
    tea.load_state_dict({k: v.to(dev) for k, v in cases.mhim_state(2, base, D=D).items()})
    for m_ in list(stu.modules()) + list(tea.modules()):
        if isinstance(m_, torch.nn.Dropout): m_.p = 0.0
    xb = cases.make_bag(3, N, D).to(dev)
    res = {}
    def teacher():
        return tea.forward_teacher(xb)
    res["teacher_ms"] = gpu_time(teacher)
    ct, sc = teacher()
    tcf = ct[0] if base == "dsmil" else ct
    if student:
        def full():
            stu.zero_grad(set_to_none=True)
            ct_, sc_ = tea.forward_teacher(xb)
            t_ = ct_[0] if base == "dsmil" else ct_
            lg, loss, _, _ = stu(xb, sc_, t_, i=0)
            lt = 0.5 * lg[0].view(1, -1) + 0.5 * lg[1].view(1, -1) if base == "dsmil" else lg
            (F.cross_entropy(lt, LABEL) + 0.5 * loss).backward()
        res["train_step_ms"] = gpu_time(full)
        from mhimk.engines import GraphedStep
        gfull = GraphedStep(lambda bag: full())
        res["train_step_graphed_ms"] = gpu_time(lambda: gfull(xb))
    from mhimk.engines import GraphedStep
    with torch.no_grad():                                   # inference replayed from a CUDA graph (input filled in place: no copy per replay)
        gt = GraphedStep(lambda bag: tea.forward_teacher(bag))
        buf = gt.buffers(xb)[0]
        res["teacher_graphed_ms"] = gpu_time(lambda: gt(buf))
    stu.eval()
    res["forward_test_ms"] = gpu_time(lambda: stu.forward_test(xb))
    with torch.no_grad():
        gf = GraphedStep(lambda bag: stu.forward_test(bag))
        buf2 = gf.buffers(xb)[0]
        res["forward_test_graphed_ms"] = gpu_time(lambda: gf(buf2))
    # informational: the oracle's torch ops on the same GPU (eager PyTorch CUDA: aten / cuBLAS / cuDNN kernels)
    cfg_ = O.MHIMConfig(**dict(cases.MHIM_KW, baseline=base, input_dim=D))
    sds_g = {k: v.to(dev) for k, v in cases.mhim_state(1, base, D=D).items()}
    sdt_g = {k: v.to(dev) for k, v in cases.mhim_state(2, base, D=D).items()}
    with torch.no_grad():
        res["eager_torch_cuda_forward_test_ms"] = gpu_time(lambda: O.mhim_forward_test(cfg_, sds_g, xb))
        res["eager_torch_cuda_teacher_ms"] = gpu_time(lambda: O.mhim_forward_teacher(cfg_, sdt_g, xb))
    if student:
        sdl_g = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sds_g.items()}
        def eager_step():
            for v in sdl_g.values():
                v.grad = None
            with torch.no_grad():
                rc_, rs_ = O.mhim_forward_teacher(cfg_, sdt_g, xb)
            t_ = rc_[0] if base == "dsmil" else rc_
            lg, loss, *_ = O.mhim_forward(cfg_, sdl_g, xb, rs_, t_, i=0, training=True)
            lt = 0.5 * lg[0].view(1, -1) + 0.5 * lg[1].view(1, -1) if base == "dsmil" else lg
            (F.cross_entropy(lt, LABEL) + 0.5 * loss).backward()
        res["eager_torch_cuda_train_step_ms"] = gpu_time(eager_step, max(2, REPS // 2))
    if CPU and do_cpu:
        cfg = O.MHIMConfig(**dict(cases.MHIM_KW, baseline=base, input_dim=D))
        sds, sdt = cases.mhim_state(1, base, D=D), cases.mhim_state(2, base, D=D)
        xc = xb.cpu()
        with torch.no_grad():
            res["cpu_forward_test_ms"] = cpu_time(lambda: O.mhim_forward_test(cfg, sds, xc), 2)
            res["cpu_teacher_ms"] = cpu_time(lambda: O.mhim_forward_teacher(cfg, sdt, xc), 2)
    out[tag] = res

def guarded(tag, fn):
    """One config failing (e.g. out of time on a busy box) must not lose the others: record the error and go on."""
    try:
        fn()
    except Exception as e:  # noqa: BLE001 - measurement script
        out[tag + "_error"] = f"{type(e).__name__}: {e}"[:300]
    print(json.dumps(out), flush=True)


guarded("cfg1", lambda: mhim_cfg("attn", 10000, 1024, "cfg1_mhim_attn_N10000_D1024"))
guarded("cfg3", lambda: mhim_cfg("dsmil", 10000, 1536, "cfg3_mhim_dsmil_N10000_D1536"))
guarded("cfg2", lambda: mhim_cfg("selfattn", 50000, 1024, "cfg2_mhim_selfattn_N50000_D1024",
                                 do_cpu=os.environ.get("CFG_CPU_BIG", "0") == "1", student=True))


def transmil():
    # plain TransMIL eval forward at N=50k
    t = M.TransMIL(1024, 2, dropout=0.0, act="relu").to(dev).eval()
    xb = cases.make_bag(5, 50000, 1024).to(dev)
    sd_g = {k: v.detach() for k, v in t.state_dict().items()}
    with torch.no_grad():
        out["transmil_eval_fwd_N50000_ms"] = gpu_time(lambda: t(xb), 5)
        from mhimk.engines import GraphedStep
        gtm = GraphedStep(lambda bag: t(bag))
        bufm = gtm.buffers(xb)[0]
        out["transmil_eval_fwd_graphed_N50000_ms"] = gpu_time(lambda: gtm(bufm), 5)
        out["transmil_eager_torch_cuda_N50000_ms"] = gpu_time(lambda: O.transmil_forward(sd_g, xb, "relu"), 3)
    # the headline module next to eager torch
    a = M.DAttention(1024, 2, dropout=0.0, act="relu").to(dev).eval()
    sda = {k: v.detach() for k, v in a.state_dict().items()}
    with torch.no_grad():
        out["abmil_N50000_fused_ms"] = gpu_time(lambda: a(xb), 20)
        out["abmil_N50000_eager_torch_cuda_ms"] = gpu_time(lambda: O.abmil_dattention(sda, xb, "relu"), 10)


guarded("transmil", transmil)


def clam():
    # f-3: CLAM_SB training step (instance branch on, dropout 0.25) at N = 10 000, eager / CUDA graph; the oracle's torch ops on the GPU beside it
    from mhimk.engines import GraphedStep
    N = 10000
    m = M.CLAM_SB(input_dim=1024, n_classes=2, dropout=0.25, act="relu").to(dev).train()
    xb = cases.make_bag(6, N, 1024)[0].to(dev)
    lab = torch.tensor([1], device=dev)
    def step(bag, label):
        m.zero_grad(set_to_none=True)
        lg, il, _ = m(bag[None], label=label, instance_eval=True)
        (F.cross_entropy(lg, label) + il).backward()
    res = {"train_step_ms": gpu_time(lambda: step(xb, lab))}
    gs = GraphedStep(step)
    res["train_step_graphed_ms"] = gpu_time(lambda: gs(xb, lab))
    m.eval()
    with torch.no_grad():
        res["eval_fwd_ms"] = gpu_time(lambda: m(xb[None]))
    sd0 = cases.clam_state(1, False)
    sdl = {k: v.to(dev).requires_grad_(v.is_floating_point()) for k, v in sd0.items()}
    arange2 = torch.arange(2)
    def eager():
        for v in sdl.values():
            v.grad = None
        lg, il, _ = O.clam_forward(sdl, xb, False, label=1)
        (F.cross_entropy(lg, lab) + il).backward()
    try:
        res["eager_torch_cuda_train_step_ms"] = gpu_time(eager, max(2, REPS // 2))
    except Exception as e:  # noqa: BLE001 - the CPU oracle builds a few index tensors on the host
        res["eager_torch_cuda_train_step_error"] = str(e)[:120]
    out["clam_sb_N10000_D1024"] = res


guarded("clam", clam)
print(json.dumps(out, indent=1))
