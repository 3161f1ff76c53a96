"""Diagnostic: one driver-shaped TransR training step, reference class on the host vs CUDA module, same tables / batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "joint-kg-recommender_b200"), os.path.join(ROOT, "baseline")):
    sys.path.insert(0, p)
import numpy as np, torch
import ref_arm
R = ref_arm.load()
import kgrec_b200 as K
L = R["loss"]
d, E, NR, B = 32, 600, 8, 256
for name, rc, oc in (("transe", R["TransE"], K.TransEModel), ("transr", R["TransR"], K.TransRModel)):
    torch.manual_seed(3)
    ref = rc(False, d, E, NR)
    torch.manual_seed(3)
    our = oc(False, d, E, NR)
    for k, v in ref.state_dict().items():
        assert torch.equal(v, our.state_dict()[k].cpu()), k
    o1 = torch.optim.Adagrad(ref.parameters(), lr=0.05, weight_decay=1e-5)
    o2 = torch.optim.Adagrad(our.parameters(), lr=0.05, weight_decay=1e-5)
    g = torch.Generator().manual_seed(0)
    for step in range(3):
        ph, pt = torch.randint(0, E, (B,), generator=g), torch.randint(0, E, (B,), generator=g)
        pr = torch.randint(0, NR, (B,), generator=g)
        nh, nt = ph.clone(), torch.randint(0, E, (B,), generator=g)
        res = []
        for m, opt, dev in ((ref, o1, "cpu"), (our, o2, "cuda")):
            a = [x.to(dev) for x in (ph, pt, pr, nh, nt)]
            opt.zero_grad()
            pos, neg = m(a[0], a[1], a[2]), m(a[3], a[4], a[2])
            loss = L.marginLoss()(pos, neg, 1.0) if dev == "cpu" else torch.sum(torch.clamp(pos - neg + 1.0, min=0))
            ent = m.ent_embeddings(torch.cat([a[0], a[1], a[3], a[4]]))
            rel = m.rel_embeddings(torch.cat([a[2], a[2]]))
            nl = lambda e: torch.sum(torch.clamp(torch.sum(e ** 2, dim=1, keepdim=True) - 1.0, min=0))
            loss = loss + nl(ent) + nl(rel)
            loss.backward()
            gn = torch.nn.utils.clip_grad_norm_(list(m.parameters()), 5.0)
            grads = {k: p.grad.detach().cpu().clone() for k, p in m.named_parameters()}
            opt.step()
            res.append((loss.item(), float(gn), grads, pos.detach().cpu(), neg.detach().cpu()))
        (l1, n1, g1, p1, q1), (l2, n2, g2, p2, q2) = res
        print(name, "step", step, "loss", l1, l2, "gradnorm", n1, n2, "pos maxrel", float(((p1 - p2).abs() / p1.abs().clamp_min(1e-6)).max()),
              "neg maxrel", float(((q1 - q2).abs() / q1.abs().clamp_min(1e-6)).max()))
        for k in g1:
            dlt = (g1[k] - g2[k]).abs().max().item()
            print("   grad", k, "max|ref|", g1[k].abs().max().item(), "max abs diff", dlt)
        for (k, a), (_, b) in zip(ref.state_dict().items(), our.state_dict().items()):
            print("   weight", k, "max abs diff", (a - b.cpu()).abs().max().item())
