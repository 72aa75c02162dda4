import sys, os, tempfile, pathlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pytest
rc = pytest.main(['tests/test_gpu_b_models.py', 'tests/test_gpu_c_step.py', '-q', '-k', 'not eval_batch'])
import torch
from oracle import patchgan_oracle as orc
from patchgan_b200 import engine as E
import tests.test_gpu_c_step as T
gk, dk, loss_type, B, steps = T.CASES['tversky']
tr, otr, _ = T.build(gk, dk, loss_type, pathlib.Path(tempfile.mkdtemp()))
tr.generator.eval(); tr.discriminator.eval()
x, y = orc.synthetic_batch(B, 1, 256, seed=99)
ref = otr.batch(x, y, train=False)
w_before = {k: p.detach().clone() for k, p in tr.generator.named_parameters()}
got = tr.batch(x, y, train=False)
print('got', got)
print('ref', ref)
xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
losses = tr.step_device(xd, yd, False)
torch.cuda.synchronize()
print('losses', losses.cpu().numpy())
for i, t in enumerate(E._KEEP):
    tf = t.float()
    nan = int(torch.isnan(tf).sum()); inf = int(torch.isinf(tf).sum())
    print(i, tuple(t.shape), t.dtype, 'nan', nan, 'inf', inf, 'absmax', float(tf[~torch.isnan(tf)].abs().max()) if nan < tf.numel() else None)
