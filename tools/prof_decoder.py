"""tools/prof_decoder.py -- run a few eager bf16 engine forwards at the DanceTrack size (target for `ncu -k regex:decoder_fused`)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from memotr_b200 import synthetic as synth
from memotr_b200.engine import FrameEngine

cfg = synth.dancetrack_cfg()
sd = synth.reference_init_state_dict(cfg, seed=0)
x = synth.frame_inputs(cfg, synth.DANCETRACK_SHAPES, 100, seed=1)
eng = FrameEngine(sd, cfg, synth.DANCETRACK_SHAPES, 100, "cuda", mode="bf16")
eng.load_frame(x["srcs"], x["masks"], x["pos"], x["tracks"]["ref_pts"], x["tracks"]["query_embed"])
for _ in range(3):
    eng.forward()
torch.cuda.synchronize()

# phase timestamps of the fused decoder kernel (clock64 per boundary, see STAMP in csrc/decoder_fused.cu)
if eng.dec_fused:
    nb = (eng.nq + 15) // 16 * (4 if eng.dec_cluster else 1)
    prof = torch.zeros(nb, eng.n_dec, 16, dtype=torch.int64, device="cuda")
    params = eng.dec_params_cl if eng.dec_cluster else eng.dec_params
    params.prof = prof.data_ptr()
    eng.forward()
    torch.cuda.synchronize()
    p = prof.cpu().double()
    names = ["query_pos", "qkv_proj", "barrier", "attention", "sa_out+ln", "ol_gemm", "gather", "ca_out+ln", "ffn+ln",
             "store", "heads"]
    d = (p[:, :, 1:12] - p[:, :, 0:11]) / 1.9e3          # ~us at 1.9 GHz
    for l in range(eng.n_dec):
        print("layer", l, {n: round(float(d[:, l, i].mean()), 1) for i, n in enumerate(names)},
              "total", round(float((p[:, l, 11] - p[:, l, 0]).mean() / 1.9e3), 1))
    print("mean per phase over layers:", {n: round(float(d[:, :, i].mean()), 1) for i, n in enumerate(names)})
    params.prof = None
