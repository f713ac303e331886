"""Timeline of ONE pipelined training step inside its CUDA graph: %globaltimer stamps (vlnce_debug_stamp) captured on
the branch each phase runs on — where the head chain and the two trunks actually sit relative to each other when they
share the GPU.  Prints the stamps of the last of a few replays, relative to the step's start.  Not the benchmark."""
import json
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.filterwarnings("ignore")
import vlnce_b200  # noqa: E402
from vlnce_b200 import lib, synth  # noqa: E402
from vlnce_b200.il_step import ILTrainerStep  # noqa: E402
from vlnce_b200.lib import C  # noqa: E402


def main():
    n = 64
    pol = synth.build_cma_policy().cuda().eval()
    vlnce_b200.AuxLosses.activate()
    step = ILTrainerStep(pol)
    obs, prev, masks, actions, weights = (
        ({k: t.cuda() for k, t in v.items()} if isinstance(v, dict) else v.cuda()) for v in synth.make_train_batch(1, 1, n))
    eng = pol.net.engine()
    eng.ensure_pool(torch.device("cuda"))
    step._prepare()
    stamps = torch.zeros(64, dtype=torch.int64, device="cuda")
    names = []

    def stamp(name):
        if name not in names:
            names.append(name)
        C.vlnce_debug_stamp(stamps.data_ptr() + 8 * names.index(name), lib.stream())

    def wrap(obj, attr, tag):
        fn = getattr(obj, attr)

        def wrapped(*a, **k):
            stamp(tag + ":begin")
            r = fn(*a, **k)
            stamp(tag + ":end")
            return r

        setattr(obj, attr, wrapped)

    wrap(pol.net.rgb_encoder, "features_into", "rgb_trunk")
    wrap(pol.net.depth_encoder, "features_into", "depth_trunk")
    wrap(eng, "_encode_instruction", "instr_encoder(fwd LSTM)")
    wrap(eng, "forward", "head_forward")
    wrap(eng, "run_backward", "head_backward")
    wrap(step, "apply_grads", "adam")
    import vlnce_b200.backward as bw

    orig_ib = bw.instruction_backward

    def ib(*a, **k):
        stamp("instr_backward(bwd LSTM):begin")
        r = orig_ib(*a, **k)
        stamp("instr_backward(bwd LSTM):end")
        return r

    bw.instruction_backward = ib
    vis = {k: obs[k] for k in ("rgb", "depth")}
    obs_nv = {k: v for k, v in obs.items() if k not in ("rgb", "depth")}
    step.prefetch_visual(vis)
    step.rotate_visual(n)

    def one():
        stamp("step:begin")
        step.pipelined_step(obs_nv, prev, masks, actions, weights, vis)
        stamp("step:end")

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        one()
        one()
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            one()
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    t = stamps.cpu().tolist()
    t0 = t[names.index("step:begin")]
    rows = sorted(((t[i] - t0) / 1e3, nm) for i, nm in enumerate(names))
    for us, nm in rows:
        print(f"{us:9.1f} us  {nm}")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({nm: us for us, nm in rows}, open("gpurun_out/timeline.json", "w"), indent=1)


if __name__ == "__main__":
    main()
