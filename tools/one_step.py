"""One small wsgan_emb training step (eager launches) for compute-sanitizer / ncu: B pairs at S x S, encoder at SE."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pcgan_b200.wsgan_emb_model import WSGANEmbModel, default_options

B, S, SE, STEPS = (int(os.environ.get(k, d)) for k, d in (("B", "2"), ("S", "32"), ("SE", "64"), ("STEPS", "1")))
torch.manual_seed(0)
opt = default_options(batchSize=B, gpu_ids=[0], fineSize=S, loadSize=S, fineSize_E=SE, which_model_netG=os.environ.get("G", "resnet_6blocks"))
m = WSGANEmbModel()
m.initialize(opt)
m.setup(opt)
for it in range(STEPS):
    m.set_input({"A": torch.rand(B, 3, S, S) * 2 - 1, "B": torch.rand(B, 3, S, S) * 2 - 1, "label": torch.randint(0, 3, (B,))})
    m.optimize_parameters()
torch.cuda.synchronize()
print("losses", {k: round(v, 5) for k, v in m.get_current_losses().items()})
