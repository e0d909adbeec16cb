"""Stand-alone probe of the tcgen05 kernels against the SIMT kernels and a CPU fp64 reference.

Usage:  python tests/tc_probe.py            # runs every case, each in its own subprocess
        python tests/tc_probe.py <case>     # run one case in-process
Each case prints one line: name, max abs error, reference scale, verdict.
A wrong descriptor shows up as a large error or a trapped kernel, never as a hang
(kernel waits are bounded; every subprocess has a timeout).
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))

CASES = {
    # name: (kind, N, H, W, Cin, Cout, k)
    'fprop_lin_128x128': ('fprop', 128, 1, 1, 128, 128, 1),
    'fprop_lin_64x2048': ('fprop', 64, 1, 1, 128, 2048, 1),
    'fprop_1x1_8x8': ('fprop', 4, 8, 8, 128, 128, 1),
    'fprop_3x3_8x8': ('fprop', 4, 8, 8, 128, 128, 3),
    'fprop_3x3_16x16': ('fprop', 3, 16, 16, 128, 128, 3),
    'fprop_3x3_32x32': ('fprop', 2, 32, 32, 128, 128, 3),
    'fprop_3x3_4x4_c256': ('fprop', 5, 4, 4, 256, 64, 3),
    'fprop_5x5_7x7': ('fprop', 3, 7, 7, 64, 128, 5),
    'dgrad_3x3_16x16': ('dgrad', 2, 16, 16, 128, 128, 3),
    'wgrad_lin': ('wgrad', 128, 1, 1, 128, 128, 1),
    'wgrad_1x1_8x8': ('wgrad', 4, 8, 8, 128, 128, 1),
    'wgrad_3x3_8x8': ('wgrad', 4, 8, 8, 128, 128, 3),
    'wgrad_3x3_32x32': ('wgrad', 2, 32, 32, 128, 256, 3),
    'wgrad_5x5_7x7': ('wgrad', 3, 7, 7, 128, 128, 5),
}


def run_case(name):
    import torch
    from ctgan_b200 import kernels as K
    kind, N, H, W, Cin, Cout, k = CASES[name]
    torch.manual_seed(0)
    dev = 'cuda'
    g = K.same_geom(N, H, W, Cin, Cout, k, 1)
    x = torch.randn(N, Cin, H, W, device=dev).to(memory_format=torch.channels_last).to(torch.bfloat16)
    dy = torch.randn(N, Cout, H, W, device=dev).to(memory_format=torch.channels_last).to(torch.bfloat16)
    w = (torch.randn(k, k, Cin, Cout, device=dev) * 0.05).contiguous()
    bias = torch.randn(Cout, device=dev)
    wq = w.to(torch.bfloat16).to(torch.float32)        # the TC path rounds the filter to bf16
    if H == 1 and W == 1:
        x, dy = x.reshape(N, Cin).contiguous(), dy.reshape(N, Cout).contiguous()
    K.config.use_tc = True
    if kind == 'fprop':
        out = K.conv_fprop(x, w, bias, g).float()
        K.config.use_tc = False
        ref = K.conv_fprop(x.float(), wq, bias, g)
    elif kind == 'dgrad':
        out = K.conv_dgrad(dy, w, g).float()
        K.config.use_tc = False
        ref = K.conv_dgrad(dy.float(), wq, g)
    else:
        out = K.conv_wgrad(x, dy, g, tuple(w.shape))
        K.config.use_tc = False
        ref = K.conv_wgrad(x.float(), dy.float(), g, tuple(w.shape))
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    # bf16 output rounding: 2^-8 relative on the largest values
    tol = scale * (1.0 / 128) if kind != 'wgrad' else scale * 1e-3
    ok = err <= tol
    print('%-22s max_err %.4e  ref_max %.4e  tol %.2e  %s' % (name, err, scale, tol, 'OK' if ok else 'MISMATCH'), flush=True)
    if not ok:
        d = (out - ref).abs()
        idx = torch.nonzero(d > tol)
        print('   mismatching elements: %d of %d; first: %s' % (idx.shape[0], d.numel(), idx[:6].tolist()), flush=True)
        flat_o, flat_r = out.flatten(), ref.flatten()
        print('   out[:8]=%s\n   ref[:8]=%s' % (flat_o[:8].tolist(), flat_r[:8].tolist()), flush=True)
    return ok


if __name__ == '__main__':
    if len(sys.argv) > 1:
        sys.exit(0 if run_case(sys.argv[1]) else 1)
    bad = 0
    for name in CASES:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=120,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            out = r.stdout.strip().splitlines()
            tail = [l for l in out if l.strip()][-6:]
            print('\n'.join(tail), flush=True)
            if r.returncode != 0:
                bad += 1
                if not any('MISMATCH' in l for l in tail):
                    print('%-22s FAILED rc=%d' % (name, r.returncode), flush=True)
        except subprocess.TimeoutExpired:
            bad += 1
            print('%-22s TIMEOUT' % name, flush=True)
    print('tc_probe: %d of %d cases failed' % (bad, len(CASES)), flush=True)
    sys.exit(1 if bad else 0)
