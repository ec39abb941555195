// fp64 entry points of the fused layer steps (see glab_layers_impl.cuh).
#define GLAB_LAYERS_F64
#include "glab_layers_impl.cuh"
