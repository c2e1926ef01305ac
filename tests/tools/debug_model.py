"""Per-parameter gradient error listing of one training step vs the CPU oracle (run on the GPU box)."""
import sys
import torch
from oracle import torch_oracle as O
from radar_depth_b200.model.models import ResNet_latefusion
from radar_depth_b200.evaluation.criteria_new import MaskedL1Loss

precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
cin = int(sys.argv[2]) if len(sys.argv) > 2 else 4
h, w = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (64, 96)
sd = O.synth_state_dict(O.latefusion_entries(cin))
inputs, target = O.synth_batch(2, h, w)
if cin == 5:
    gen = torch.Generator().manual_seed(99)
    inputs = torch.cat((inputs, torch.rand(2, 1, h, w, generator=gen) * 40), dim=1)
ref = O.train_step(sd, inputs, target, "latefusion", dtype=torch.float64)
m = ResNet_latefusion(18, "upproj", (h, w), cin, pretrained=False)
m.load_state_dict(sd, strict=True)
m = m.cuda().train()
m.precision = precision
pred = m(inputs.cuda())
loss = MaskedL1Loss()(pred, target.cuda())
loss.backward()
torch.cuda.synchronize()
print("pred rel", float((pred.double().cpu() - ref["pred"]).norm() / ref["pred"].norm()), "loss", float(loss), float(ref["loss"]))
for k, p in m.named_parameters():
    g, r = p.grad.double().cpu(), ref["grads"][k]
    rel = float((g - r).norm() / (r.norm() + 1e-30))
    cos = float((g * r).sum() / (g.norm() * r.norm() + 1e-30))
    flag = "" if rel < 2e-2 else "   <<<<"
    print(f"{k:55s} rel {rel:9.3e} cos {cos:+.4f} |g| {float(g.norm()):9.3e} |ref| {float(r.norm()):9.3e}{flag}")
named = dict(m.named_parameters())
print("bn1.bias.grad[:4]", named["bn1.bias"].grad[:4].cpu().numpy(), "ref", ref["grads"]["bn1.bias"][:4].numpy())
print("conv1.weight.grad[0,0,0]", named["conv1.weight"].grad[0, 0, 0].cpu().numpy(), "ref", ref["grads"]["conv1.weight"][0, 0, 0].numpy())
# second, independent oracle evaluation: same function again
ref2 = O.train_step(sd, inputs, target, "latefusion", dtype=torch.float64)
print("oracle repeat: bn1.bias", float((ref2["grads"]["bn1.bias"] - ref["grads"]["bn1.bias"]).norm()), "ref2[:4]", ref2["grads"]["bn1.bias"][:4].numpy())
ref3 = O.train_step(sd, inputs, target, "latefusion", dtype=torch.float32)
print("oracle fp32: bn1.bias[:4]", ref3["grads"]["bn1.bias"][:4].numpy())
