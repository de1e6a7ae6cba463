import sys, torch, time
sys.path.insert(0, "/root/repo")
from cddmsl_b200 import synth
from cddmsl_b200.layers import nms
for m in (16000, 24576, 64000, 131072, 262144):
    g = synth.generator(9)
    boxes, scores, _ = synth.make_nms_inputs(m, 1024, 2048, g, tie_frac=0.0)
    bd, sd = boxes.cuda(), scores.cuda()
    for _ in range(2): k = nms(bd, sd, 0.7)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): k = nms(bd, sd, 0.7)
    e1.record(); torch.cuda.synchronize()
    print(m, "boxes:", round(e0.elapsed_time(e1) / 3, 3), "ms, kept", len(k))
