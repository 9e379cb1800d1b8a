// The one place the Rust host needs libtorch's C++ API: tch-rs does not expose the current CUDA stream, and work must
// be issued on it (not on the legacy stream) so that it is ordered with the tensors' producers and consumers.
// Compiled by build.rs with the LIBTORCH include path the reference's torch-sys build already requires
// (g++ -I$LIBTORCH/include -c torch_stream_shim.cpp; links against c10_cuda).  Not part of libtchgeo_cuda.so: the Python
// mirror takes the stream from torch.cuda.current_stream() instead.
#include <c10/cuda/CUDAStream.h>

extern "C" void* tchgeo_torch_current_stream(int device_index) {
  return (void*)c10::cuda::getCurrentCUDAStream((c10::DeviceIndex)device_index).stream();
}
