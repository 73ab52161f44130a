// pil_unity.cu -- single-translation-unit build of libpil.so for the development instrumentation that keeps
// __device__ globals (-DPIL_BOUNDS: tools/bounds_check.py, -DPIL_TIMELINE: tools/timeline.py).  Release builds
// compile the pieces separately and in parallel (physics_informed_image_segmentation_b200/_lib.py).
#include "pil_fwd.cu"
#include "pil_point.cu"
#include "pil_bwd.cu"
#include "pil_api.cu"
#include "pil_tail.cu"
