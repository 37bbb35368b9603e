"""Times forward / backward / optimiser of one training step (train_nerf_regtr.py:171-239) at a given resolution."""
import sys
import os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dreg_nerf_b200 as pkg
from oracle.make_goldens import training_loss   # a fixed scalar of the outputs (test infrastructure, not the product)

res = int(sys.argv[1]) if len(sys.argv) > 1 else 128
precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = pkg.NeRFRegTr(precision=precision)
model.load_state_dict(pkg.synthetic.seeded_state_dict(model, seed=0, attn_gain=4.0))
model = model.to(dev).train(True)
model.correspondence_decoder.q_norm.requires_grad_(False)
opt = pkg.FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, weight_decay=1e-4, max_grad_norm=0.1)
pairs = [pkg.synthetic.to_device(pkg.synthetic.make_pair(res=res, pair_id=i), dev) for i in range(3)]
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
acc = [0.0, 0.0, 0.0]
for it in range(steps + 2):
    opt.zero_grad(set_to_none=True)
    ev[0].record()
    out = model(dict(pairs[it % 3]))
    loss = training_loss(out)
    ev[1].record()
    loss.backward()
    ev[2].record()
    opt.step()
    ev[3].record()
    torch.cuda.synchronize()
    if it >= 2:
        for j in range(3):
            acc[j] += ev[j].elapsed_time(ev[j + 1])
    print("step %d loss %.5f tokens %s fwd %.2f bwd %.2f opt %.2f ms" % (it, float(loss), model.last_token_counts,
          ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])), flush=True)
print("mean over %d steps: fwd %.2f bwd %.2f opt %.2f ms  (%s, %d^3)" % (steps, acc[0] / steps, acc[1] / steps, acc[2] / steps, precision, res))
print("launches", model.launch_count())
