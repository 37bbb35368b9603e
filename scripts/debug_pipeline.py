"""Debug aid: gradients of 5 pairs accumulated sequentially vs through PairPipeline under several locking regimes."""
import sys
import threading

import torch

sys.path.insert(0, ".")
import dreg_nerf_b200 as pkg
from oracle.make_goldens import training_loss

cuda = torch.device("cuda:0")
torch.manual_seed(0)
model = pkg.NeRFRegTr(precision="fp32")
model.load_state_dict(pkg.synthetic.seeded_state_dict(model, seed=0, attn_gain=4.0))
model = model.to(cuda).train(len(sys.argv) > 1 and sys.argv[1] == 'train')
model.correspondence_decoder.q_norm.requires_grad_(False)
pairs = [pkg.synthetic.to_device(pkg.synthetic.make_pair(res=32, pair_id=10 + i), cuda) for i in range(5)]
named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
lock = threading.Lock()
mode = {"lock": "none"}


def one(data):
    if mode["lock"] == "all":
        with lock:
            loss = training_loss(model(dict(data)))
            (loss / len(pairs)).backward()
            return loss.detach()
    loss = training_loss(model(dict(data)))
    if mode["lock"] == "bwd":
        with lock:
            (loss / len(pairs)).backward()
    else:
        (loss / len(pairs)).backward()
    return loss.detach()


def grads():
    torch.cuda.synchronize()
    return {n: p.grad.detach().clone() for n, p in named if p.grad is not None}


def diff(a, b, tag):
    rows = []
    for n in a:
        rows.append((((a[n] - b[n]).abs().max() / a[n].abs().max().clamp_min(1e-30)).item(), n))
    rows.sort(reverse=True)
    print(tag, "worst:", ["%.1e %s" % r for r in rows[:6]], "tensors > 1e-5:", sum(1 for r in rows if r[0] > 1e-5), flush=True)


model.zero_grad(set_to_none=True)
for p in pairs:
    one(p)
want = grads()
model.zero_grad(set_to_none=True)
for p in pairs:
    one(p)
diff(want, grads(), "sequential twice")
pipe = pkg.PairPipeline(cuda, streams=3)
for lk in ("all", "bwd", "none", "none"):
    mode["lock"] = lk
    model.zero_grad(set_to_none=True)
    pipe.map(one, pairs)
    diff(want, grads(), "pipeline lock=%s" % lk)
# one pair at a time through a NON-zero slot (a secondary engine alone)
mode["lock"] = "none"
model.zero_grad(set_to_none=True)
for p in pairs:
    pipe.map(one, [p, p])          # both slots run the same pair -> 2x the gradient of p
g2 = grads()
diff(want, {k: v / 2 for k, v in g2.items()}, "pairs twice on 2 slots / 2")
pipe.close()
