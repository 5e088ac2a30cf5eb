// Thin PyTorch C++ extension over the C ABI (include/bilby_b200.h): the north star's "thin PyTorch C++/CUDA extension".
// TORCH_LIBRARY ops take tensors, run on the current CUDA stream of the tensor's device and return tensors, so Python
// callers (bilby_b200/gw/likelihood.py) and TorchScript / C++ hosts need no data_ptr() / stream plumbing.  The ops are
// one-to-one with the batched entry points that replace GravitationalWaveTransient.log_likelihood_ratio / calculate_snrs
// (bilby/gw/likelihood/base.py:419-446, 260-354); `handle` is the bb_handle* as an integer.  No arithmetic lives here.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include "../../include/bilby_b200.h"

namespace {

bb_handle* as_handle(int64_t h) {
    TORCH_CHECK(h != 0, "bilby_b200: null handle");
    return reinterpret_cast<bb_handle*>(h);
}

void check_rows(const at::Tensor& rows) {
    TORCH_CHECK(rows.is_cuda(), "bilby_b200: parameter rows must be a CUDA tensor (there is no CPU path)");
    TORCH_CHECK(rows.scalar_type() == at::kDouble && rows.dim() == 2 && rows.size(1) == BB_NPARAM && rows.is_contiguous(),
                "bilby_b200: parameter rows must be a contiguous float64 tensor of shape [n, ", BB_NPARAM, "]");
}

void* stream_of(const at::Tensor& t) { return c10::cuda::getCurrentCUDAStream(t.device().index()).stream(); }

void check_rc(int rc) { TORCH_CHECK(rc == 0, "bilby_b200: ", bb_last_error()); }

at::Tensor log_likelihood_ratio(int64_t handle, const at::Tensor& rows) {
    check_rows(rows);
    at::Tensor out = at::empty({rows.size(0)}, rows.options());
    check_rc(bb_log_likelihood_ratio_device(as_handle(handle), rows.data_ptr<double>(), rows.size(0), out.data_ptr<double>(),
                                            stream_of(rows)));
    return out;
}

at::Tensor log_likelihood_ratio_cal(int64_t handle, const at::Tensor& rows, const at::Tensor& cal) {
    check_rows(rows);
    TORCH_CHECK(cal.is_cuda() && cal.scalar_type() == at::kDouble && cal.is_contiguous() && cal.size(0) == rows.size(0),
                "bilby_b200: calibration parameters must be a contiguous CUDA float64 tensor [n, n_det, 2, n_points]");
    at::Tensor out = at::empty({rows.size(0)}, rows.options());
    check_rc(bb_log_likelihood_ratio_cal_device(as_handle(handle), rows.data_ptr<double>(), cal.data_ptr<double>(),
                                                rows.size(0), out.data_ptr<double>(), stream_of(rows)));
    return out;
}

at::Tensor inner_products(int64_t handle, const at::Tensor& rows, int64_t n_det) {
    check_rows(rows);
    at::Tensor out = at::empty({rows.size(0), n_det, 3}, rows.options());
    check_rc(bb_inner_products_device(as_handle(handle), rows.data_ptr<double>(), rows.size(0), out.data_ptr<double>(),
                                      stream_of(rows)));
    return out;
}

at::Tensor likelihood_from_inner_products(int64_t handle, const at::Tensor& rows, const at::Tensor& snrs) {
    check_rows(rows);
    TORCH_CHECK(snrs.is_cuda() && snrs.scalar_type() == at::kDouble && snrs.is_contiguous() && snrs.size(0) == rows.size(0),
                "bilby_b200: inner products must be a contiguous CUDA float64 tensor [n, n_det, 3]");
    at::Tensor out = at::empty({rows.size(0)}, rows.options());
    check_rc(bb_likelihood_from_inner_products_device(as_handle(handle), rows.data_ptr<double>(), snrs.data_ptr<double>(),
                                                      rows.size(0), out.data_ptr<double>(), stream_of(rows)));
    return out;
}

}  // namespace

TORCH_LIBRARY(bilby_b200, m) {
    m.def("log_likelihood_ratio(int handle, Tensor rows) -> Tensor", &log_likelihood_ratio);
    m.def("log_likelihood_ratio_cal(int handle, Tensor rows, Tensor cal) -> Tensor", &log_likelihood_ratio_cal);
    m.def("inner_products(int handle, Tensor rows, int n_det) -> Tensor", &inner_products);
    m.def("likelihood_from_inner_products(int handle, Tensor rows, Tensor snrs) -> Tensor", &likelihood_from_inner_products);
}
