// kernels of family REC (one translation unit per family so they compile in parallel)
#include "train_dev.cuh"
namespace kgrec {
KGREC_INSTANTIATE_FAMILY(FAM_REC)
}
