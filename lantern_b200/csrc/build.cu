// placeholder, replaced below
#include "engine.h"
namespace lb200 {
void build_pending(Index& idx) { (void)idx; throw CudaError("GPU build: not implemented yet"); }
}
