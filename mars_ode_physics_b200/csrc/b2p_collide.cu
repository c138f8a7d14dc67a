// b2p_collide.cu -- device broadphase + narrowphase (placeholder until the collision kernels land).
#include <cuda_runtime.h>
#include "b2p_dev.h"
extern "C" cudaError_t b2p_launch_collide(const CollideParams *P, cudaStream_t stream)
{
    (void)P; (void)stream;
    return cudaErrorNotSupported;
}
