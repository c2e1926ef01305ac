"""Developer script: run the folded inference program repeatedly and compare its buffers with the eval program's."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import torch_oracle as O
from radar_depth_b200.model.models import ResNet_latefusion

H, W = (352, 1216) if len(sys.argv) < 2 else (int(sys.argv[1]), int(sys.argv[2]))
inputs, _ = O.synth_batch(1, H, W)
x = inputs.cuda()
m = ResNet_latefusion(18, "upproj", (H, W), 4, pretrained=False)
m.load_state_dict(O.synth_state_dict(O.latefusion_entries(4)), strict=True)
m = m.cuda().eval()
m.precision = "bf16"
eng = m._get_engine()
eng.use_graphs = False
rel = lambda a, b: float((a.float() - b.float()).norm() / (b.float().norm() + 1e-30))
with torch.no_grad():
    ref = eng.forward(x, False, inference=False).clone()
    # activated tensors of the eval program, computed from its raw buffers
    want = {}
    for blks in eng.blocks_all:
        for Bk in blks:
            b1 = Bk["b1"]
            want[Bk["pfx"] + ".a1"] = torch.relu(Bk["z1"].float() * b1.scale + b1.shift)
    want["zc2"] = eng.dbg["zc2"].float() * eng.bc2.scale + eng.bc2.shift
    for L in eng.dec:
        want[L["pfx"] + ".out"] = L["out"].float().clone()
    outs = []
    for it in range(4):
        outs.append(eng.forward(x, False, inference=True).clone())
        torch.cuda.synchronize()
        print(f"run {it}: folded vs eval program {rel(outs[-1], ref):.4e}   identical to run 0: {torch.equal(outs[-1], outs[0])}")
    for blks in eng.blocks_all:
        for Bk in blks:
            print(f"  {Bk['pfx']:20s} a1 {rel(Bk['z1'], want[Bk['pfx'] + '.a1']):.3e}")
    print(f"  zc2 {rel(eng.dbg['zc2'], want['zc2']):.3e}")
    for L in eng.dec:
        print(f"  {L['pfx']:20s} out {rel(L['out'], want[L['pfx'] + '.out']):.3e}")
