// ORACLE BUILD STUB (test infrastructure): pointnet2's cuda_utils.h includes <ATen/ATen.h> only to reach
// at::cuda::getCurrentCUDAStream(); this stub lets sampling_gpu.cu compile with plain nvcc (no torch headers).
#pragma once
