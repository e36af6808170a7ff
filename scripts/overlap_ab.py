"""Launch B of chunk i (phase space + weight + reduction, FP64-bound) on a second stream WHILE launch A of chunk i+1
(flow sample, HBM-bound) runs: u / log q double-buffered. A normally fills every SM (8 CTAs x 256 threads x 32 registers =
the whole register file), so B only fits if A leaves room (MF_CHAINA_CTAS) and B is launched with a small footprint
(MF_CHAINB_THREADS=128, MF_CHAINB_CTAS). Prints ms per step of the serial and the overlapped schedule per variant."""
import os, sys, subprocess, json
VARIANTS = [dict(), dict(MF_CHAINB_THREADS="96", MF_CHAINB_CTAS="1", OV_PRIO="1"), dict(MF_CHAINB_THREADS="96", MF_CHAINB_CTAS="2", OV_PRIO="1"),
            dict(MF_CHAINB_THREADS="96", MF_CHAINB_CTAS="1", OV_PRIO="1", MF_CHAINA_CTAS="7"),
            dict(MF_CHAINB_THREADS="128", MF_CHAINB_CTAS="1", OV_PRIO="1"), dict(MF_CHAINB_THREADS="96", MF_CHAINB_CTAS="1"),
            dict(MF_CHAINA_CTAS="7", MF_CHAINB_THREADS="128", MF_CHAINB_CTAS="1"), dict(MF_CHAINA_CTAS="6", MF_CHAINB_THREADS="128", MF_CHAINB_CTAS="1")]
if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from memflow_b200 import estimator as ES
    dev = torch.device("cuda:0")
    N = 1 << 22
    phi = torch.randn(2, N, 7, 29, device=dev, dtype=torch.float32)
    smp = ES.ImportanceSampler([125.25, 172.5, 172.5], 13000.0, seed=1, device=dev)
    bufs = [(torch.empty(N, 7, dtype=torch.float64, device=dev), torch.empty(N, dtype=torch.float64, device=dev)) for _ in range(2)]
    acc = ES.new_accumulator(dev)
    # OV_PRIO=1: launch B's stream gets the higher priority, so that its (few, small) CTAs take the slots that launch A's
    # retiring CTAs free before A's next CTAs do
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(priority=-1 if os.environ.get("OV_PRIO") else 0)
    K = 40
    def serial():
        for i in range(K):
            u, lq = bufs[0]
            smp.sample_flow(phi, event_offset=i * N, u=u, logq=lq); smp.weigh(u, lq, acc=acc)
    def overlapped():
        doneA = [None, None]; doneB = [None, None]
        for i in range(K):
            u, lq = bufs[i & 1]
            with torch.cuda.stream(s1):
                if doneB[i & 1] is not None: s1.wait_event(doneB[i & 1])      # B of chunk i-2 has read this buffer
                smp.sample_flow(phi, event_offset=i * N, u=u, logq=lq)
                doneA[i & 1] = torch.cuda.Event(); doneA[i & 1].record(s1)
            with torch.cuda.stream(s2):
                s2.wait_event(doneA[i & 1])
                smp.weigh(u, lq, acc=acc)
                doneB[i & 1] = torch.cuda.Event(); doneB[i & 1].record(s2)
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    res = {}
    for name, f in (("serial", serial), ("overlapped", overlapped)):
        f(); torch.cuda.synchronize()
        acc.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        res[name] = a.elapsed_time(b) / K
        res[name + "_acc"] = acc.tolist()
    print(json.dumps(res))
else:
    for v in VARIANTS:
        outs = []
        for rnd in range(2):
            out = subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, **v), capture_output=True, text=True)
            try:
                outs.append(json.loads(out.stdout.strip().splitlines()[-1]))
            except Exception:
                print(v, "FAILED", out.stderr[-400:]); break
        if outs:
            print(v, "serial %.3f ms  overlapped %.3f ms   (runs: %s)  same sums: %s" % (
                min(o["serial"] for o in outs), min(o["overlapped"] for o in outs),
                ", ".join("%.3f/%.3f" % (o["serial"], o["overlapped"]) for o in outs),
                outs[0]["serial_acc"][2] == outs[0]["overlapped_acc"][2]))
