// model_const.cuh — constant-memory slots for substitution models (kernel translation units only).
// Each kernel translation unit owns a copy; the host fills the copy of the unit whose kernel it is about
// to launch through KernelTable::upload_model (kernel_api.hpp).
#pragma once

namespace {
__constant__ double c_model[MODEL_SLOT * MODEL_SLOTS];
}  // namespace
