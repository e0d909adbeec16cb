// optim.cu -- fused multi-tensor Adam with TensorFlow semantics on one flat float buffer.
//
// Replaces the per-variable ApplyAdam kernels behind tf.train.AdamOptimizer
// (TG/CT_gan_mnist.py:168-177, TG/CT_gan_cifar.py:153-154, TG/CT_gan_cifar_resnet.py:333-338):
//   m = b1*m + (1-b1)*g ; v = b2*v + (1-b2)*g^2 ; p -= lr_t * m / (sqrt(v) + eps)
// HBM-bound: 28 B/param (read p,g,m,v; write p,m,v), float4 accesses.
#include "common.cuh"

namespace ctgan {

template <int V>
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t nvec, float lr_t, float b1, float b2, float eps,
                            float gscale, const float* __restrict__ lr_t_dev) {
    ctgan::pdl_entry();
    if (lr_t_dev) lr_t = lr_t_dev[0];
    struct alignas(4 * V) F { float a[V]; };
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        F pp = reinterpret_cast<F*>(p)[i], gg = reinterpret_cast<const F*>(g)[i];
        F mm = reinterpret_cast<F*>(m)[i], vv = reinterpret_cast<F*>(v)[i];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            float gr = gg.a[j] * gscale;
            mm.a[j] = b1 * mm.a[j] + (1.f - b1) * gr;
            vv.a[j] = b2 * vv.a[j] + (1.f - b2) * gr * gr;
            pp.a[j] -= lr_t * mm.a[j] / (sqrtf(vv.a[j]) + eps);
        }
        reinterpret_cast<F*>(p)[i] = pp;
        reinterpret_cast<F*>(m)[i] = mm;
        reinterpret_cast<F*>(v)[i] = vv;
    }
}

}  // namespace ctgan

using namespace ctgan;

extern "C" int ctgan_adam_step(float* p, const float* g, float* m, float* v, int64_t n,
                               float lr_t, float beta1, float beta2, float eps, float grad_scale,
                               const float* lr_t_dev, void* stream) {
    CTGAN_REQUIRE(p && g && m && v, CTGAN_ERR_BAD_DESC, "adam_step: null pointer");
    if (n <= 0) return 0;
    cudaStream_t st = as_stream(stream);
    auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (n % 4 == 0 && al(p) && al(g) && al(m) && al(v)) {
        int64_t nv = n / 4;
        CTGAN_LAUNCH((adam_kernel<4>), elementwise_grid(nv, 256), 256, 0, st, p, g, m, v, nv, lr_t, beta1, beta2, eps, grad_scale, lr_t_dev);
    } else {
        CTGAN_LAUNCH((adam_kernel<1>), elementwise_grid(n, 256), 256, 0, st, p, g, m, v, n, lr_t, beta1, beta2, eps, grad_scale, lr_t_dev);
    }
    CTGAN_CHECK_LAUNCH("adam_step");
    return 0;
}
