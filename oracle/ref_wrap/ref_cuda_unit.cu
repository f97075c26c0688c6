// TEST / BASELINE INFRASTRUCTURE — NOT PRODUCT CODE.
//
// Compiles the REFERENCE's own CUDA op from the sources where they lie under /root/reference
// (no copy of reference sources enters this repository).  The only incompatibility between the
// reference and torch 2.11 is that it passes `value.type()` (a DeprecatedTypeProperties) to
// AT_DISPATCH_FLOATING_TYPES (ms_deform_attn_cuda.cu:69,139), which now wants a ScalarType.
// This unit re-defines the macro to accept both and then includes the reference translation
// unit unchanged.  REF_OPS_SRC is passed by oracle/build_ref.py as an include directory.
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>

namespace bm2f_ref_compat {
inline at::ScalarType scalar_type_of(const at::DeprecatedTypeProperties &t) { return t.scalarType(); }
inline at::ScalarType scalar_type_of(at::ScalarType t) { return t; }
}  // namespace bm2f_ref_compat

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
    AT_DISPATCH_SWITCH(bm2f_ref_compat::scalar_type_of(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))

#include "cuda/ms_deform_attn_cuda.cu"
