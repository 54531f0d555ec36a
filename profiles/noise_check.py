import sys, os, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import test_parity_gpu as T
from oracle.fixtures import random_batch, random_state
from oracle.sit_oracle import ArchSpec
from reed_b200.image.loss import SILoss
from reed_b200.image.trainer import ReedTrainer
DEV = "cuda"
spec = ArchSpec(input_size=16, hidden_size=128, decoder_hidden_size=128, depth=2, num_heads=2, encoder_depth=1, z_dims=[64], projector_dim=128)
sd = random_state(spec, 11)
batches = [random_batch(spec, 4, 40 + i) for i in range(5)]
to_dev = lambda d: (d["x"].to(DEV), d["y"].to(DEV), [z.to(DEV) for z in d["zs"]])
def run(graphed):
    torch.manual_seed(123); torch.cuda.manual_seed(123)
    model = T._build(spec, sd, "bf16").train()
    tr = ReedTrainer(model, SILoss(enc_names=["dinov2"], loss_weights={"dinov2": 1.0}), precision="bf16")
    losses = []
    if graphed:
        tr.capture(*to_dev(batches[0]), warmup=2)
        for i in range(2, 5):
            loss, _ = tr.train_step_graphed(*to_dev(batches[i]), diffusion_decay=0.5 + 0.1 * i, repa_decay=0.9)
            losses.append(float(loss))
    else:
        for i in (0, 0, 2, 3, 4):
            decay = (1.0, 1.0) if i == 0 else (0.5 + 0.1 * i, 0.9)
            loss, _ = tr.train_step(*to_dev(batches[i]), diffusion_decay=decay[0], repa_decay=decay[1])
            losses.append(float(loss))
        losses = losses[2:]
    return losses
for tag, g in (("eager", False), ("eager", False), ("eager", False), ("graph", True), ("graph", True), ("graph", True)):
    print(tag, ["%.9f" % l for l in run(g)])
