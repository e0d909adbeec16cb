"""ncu target: one batched generator forward (fakes of 5 critic steps) + one generator step, eager, batch 64."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import numpy as np, torch
import ctgan_b200.gan_cifar_resnet as R
B = 64
np.random.seed(1234)
tr = R.Trainer(device='cuda', seed=1234, act_dtype=torch.bfloat16, batch_size=B)
rs = np.random.RandomState(0)
y = torch.from_numpy(rs.randint(0, 10, (5 * B,)).astype('int32')).cuda()
for i in range(3):
    if i == 2:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    tr.generate_fakes(y)
    if i == 2:
        torch.cuda.synchronize(); print('MARK pregen done', flush=True)
    tr.gen_step()
torch.cuda.synchronize(); torch.cuda.profiler.stop()
