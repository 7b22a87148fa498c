// Opt-in instantiations of multibox_loss_kernel (loss_kernel.cuh), reached through ssdh_multibox_loss_ex only:
//   kModeForce           best-prior-per-ground-truth forcing (north_star extension of SSD._match, reference ssd.py:231-250)
//   kModeExact           libdevice exp / log / IEEE division + the cross-entropy override hook of the selection tests
//   kModeForce|kModeExact
// Kept out of loss.cu so that they compile in parallel with, and never perturb, the production kernel.
#include "loss_kernel.cuh"

namespace ssdh {

int launch_loss_extension(int mode, const LossParams& p, const LossShape& s, cudaStream_t st) {
  switch (mode) {
    case kModeForce: return launch_loss_mode<kModeForce>(p, s, st);
    case kModeExact: return launch_loss_mode<kModeExact>(p, s, st);
    case kModeForce | kModeExact: return launch_loss_mode<kModeForce | kModeExact>(p, s, st);
    default: set_error("ssdh_multibox_loss_ex: unsupported mode %d", mode); return SSDH_E_ARG;
  }
}

}  // namespace ssdh
