// C ABI plumbing: handle lifetime, error text, defaults.  The compute entry points live beside their
// kernels (matrix_kernels.cu, qphb_kernel.cu); include/hybdrt_b200.h is the contract.
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace hdrt {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace hdrt

extern "C" int hdrt_version(void) { return 100; }

extern "C" const char* hdrt_last_error(void) { return hdrt::g_err; }

extern "C" int hdrt_create(hdrt_handle** out, int device) {
    if (!out) { hdrt::set_error("hdrt_create: null out pointer"); return HDRT_ERR_ARG; }
    int count = 0;
    HDRT_CUDA_CHECK(cudaGetDeviceCount(&count));
    if (device < 0 || device >= count) { hdrt::set_error("hdrt_create: device %d of %d", device, count); return HDRT_ERR_ARG; }
    HDRT_CUDA_CHECK(cudaSetDevice(device));
    hdrt_handle* h = new hdrt_handle();
    h->device = device;
    cudaDeviceProp prop;
    HDRT_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    h->sm_count = prop.multiProcessorCount;
    h->smem_per_sm = prop.sharedMemPerMultiprocessor;
    h->smem_reserved_per_cta = prop.reservedSharedMemPerBlock;
    h->regs_per_sm = prop.regsPerMultiprocessor;
    h->threads_per_sm = prop.maxThreadsPerMultiProcessor;
    h->launches = 0;
    h->mu = new std::mutex();
    for (int i = 0; i < kLaunchSlots; ++i) {
        hdrt_launch_slot& sl = h->slots[i];
        sl.used = false;
        HDRT_CUDA_CHECK(cudaMalloc(&sl.work_counter, sizeof(int)));
        HDRT_CUDA_CHECK(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    }
    *out = h;
    return HDRT_OK;
}

extern "C" int hdrt_destroy(hdrt_handle* h) {
    if (!h) return HDRT_OK;
    cudaSetDevice(h->device);
    for (int i = 0; i < kLaunchSlots; ++i) {
        cudaFree(h->slots[i].work_counter);
        cudaEventDestroy(h->slots[i].done);
    }
    delete static_cast<std::mutex*>(h->mu);
    delete h;
    return HDRT_OK;
}

extern "C" int hdrt_sm_count(const hdrt_handle* h) { return h ? h->sm_count : 0; }

extern "C" void hdrt_default_hypers(hdrt_hypers* hyp) {
    if (!hyp) return;
    memset(hyp, 0, sizeof(*hyp));
    const double dw[3] = {1.5, 1.0, 0.5}, sig[3] = {1.0, 1000.0, 1000.0}, sa[3] = {5.0, 10.0, 25.0};
    const double ra[3] = {0.15, 0.2, 0.25}, ddw[3] = {0.5, 1.0, 0.5};
    for (int k = 0; k < 3; ++k) {
        hyp->derivative_weights[k] = dw[k]; hyp->sigma_ds[k] = sig[k]; hyp->s_alpha[k] = sa[k]; hyp->s_0[k] = 1.0;
        hyp->rho_alpha[k] = ra[k]; hyp->rho_0[k] = 1.0;
        hyp->dop_derivative_weights[k] = ddw[k]; hyp->dop_sigma_ds[k] = sig[k]; hyp->dop_s_alpha[k] = sa[k];
        hyp->dop_s_0[k] = 1.0; hyp->dop_rho_alpha[k] = ra[k]; hyp->dop_rho_0[k] = 1.0;
    }
    hyp->l2_lambda_0 = 142.0;
    hyp->dop_l2_lambda_0 = 10.0;
    hyp->iw_l1_lambda_0 = 1e-4;
    hyp->iw_l2_lambda_0 = 1e-4;
    hyp->iw_alpha = 0.0;
    hyp->iw_beta = 0.0;
    hyp->has_iw_prior = 0;
    hyp->xtol = 1e-2;
    hyp->max_iter = 50;
    hyp->weight_factor = 1.0;
    hyp->chrono_weight_factor = 1.0;
    hyp->eis_weight_factor = 1.0;
}
