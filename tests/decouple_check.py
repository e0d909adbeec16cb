"""kernels.config.decouple_gp on the GPU: one BF16 critic step of CT_gan_cifar_resnet.py against the oracle (conditioned mode),
with the stacked pass's backward issued before the gradient-penalty pass.  Prints the report; not collected by pytest."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import ctgan_b200.kernels as K
from tests import parity

K.config.decouple_gp = True
tr, om = parity.build_pair('resnet', 'cuda', torch.bfloat16, 16, oracle_dtype=torch.float32)
parity.perturb_params(tr, om)
rep = parity.critic_parity('resnet', tr, om, parity.make_inputs('resnet', 16, 11), conditioned=True, floor_frac=1e-2)
print('decoupled bf16 critic parity:', parity.format_report(rep, 6), 'gradall %.2e' % rep['gradall'])
