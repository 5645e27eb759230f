// The monomorphised scalar programs (sorted kinds of the scalar constraints), one list for the two translation units
// that instantiate kernels over it: sfgpu_scalar.cu (rows-resident scoring) and sfgpu_scalar_step.cu (generated step +
// finish). WIDE = int64 program only, BOTH = int64 and int32 forms. Kinds ascending; UC (uni without column / mask)
// sorts last. The tuples of the reference's scalar examples (graph colouring, n-queens, job shop) and their prefixes.
#pragma once
#define SFGPU_SPEC_U_ SFGPU_K_UNI
#define SFGPU_SPEC_C_ SFGPU_K_PAIR_CSR_EQUAL
#define SFGPU_SPEC_K_ SFGPU_K_PAIR_KEY_EQUAL
#define SFGPU_SPEC_G_ SFGPU_K_GROUP
#define SFGPU_SPEC_UC SPEC_K_UNI_CONST
#define SFGPU_SPEC_TUPLES(WIDE, BOTH)                                                                             \
  WIDE(SFGPU_SPEC_U_, 0, 0, 0), BOTH(SFGPU_SPEC_UC, 0, 0, 0),                               /* unassigned only */ \
  WIDE(SFGPU_SPEC_U_, SFGPU_SPEC_C_, 0, 0), BOTH(SFGPU_SPEC_C_, SFGPU_SPEC_UC, 0, 0),       /* graph colouring */ \
  WIDE(SFGPU_SPEC_U_, SFGPU_SPEC_K_, 0, 0), BOTH(SFGPU_SPEC_K_, SFGPU_SPEC_UC, 0, 0),                             \
  WIDE(SFGPU_SPEC_U_, SFGPU_SPEC_G_, 0, 0), BOTH(SFGPU_SPEC_G_, SFGPU_SPEC_UC, 0, 0),                             \
  WIDE(SFGPU_SPEC_U_, SFGPU_SPEC_C_, SFGPU_SPEC_G_, 0), BOTH(SFGPU_SPEC_C_, SFGPU_SPEC_G_, SFGPU_SPEC_UC, 0),     \
  WIDE(SFGPU_SPEC_U_, SFGPU_SPEC_K_, SFGPU_SPEC_G_, 0), BOTH(SFGPU_SPEC_K_, SFGPU_SPEC_G_, SFGPU_SPEC_UC, 0),     /* job shop */ \
  WIDE(SFGPU_SPEC_U_, SFGPU_SPEC_K_, SFGPU_SPEC_K_, 0), BOTH(SFGPU_SPEC_K_, SFGPU_SPEC_K_, SFGPU_SPEC_UC, 0),     \
  WIDE(SFGPU_SPEC_U_, SFGPU_SPEC_K_, SFGPU_SPEC_K_, SFGPU_SPEC_K_), BOTH(SFGPU_SPEC_K_, SFGPU_SPEC_K_, SFGPU_SPEC_K_, SFGPU_SPEC_UC), /* n-queens */ \
  WIDE(SFGPU_SPEC_U_, SFGPU_SPEC_K_, SFGPU_SPEC_G_, SFGPU_SPEC_G_), BOTH(SFGPU_SPEC_K_, SFGPU_SPEC_G_, SFGPU_SPEC_G_, SFGPU_SPEC_UC), \
  WIDE(SFGPU_SPEC_U_, SFGPU_SPEC_U_, SFGPU_SPEC_K_, SFGPU_SPEC_G_), WIDE(SFGPU_SPEC_U_, SFGPU_SPEC_K_, SFGPU_SPEC_G_, SFGPU_SPEC_UC)
