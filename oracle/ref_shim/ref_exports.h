// ref_exports.h — exports the translation unit's harness as extern "C" ref_<VARIANT>(RefArgs*).
#define REF_CAT2(a, b) a##b
#define REF_CAT(a, b) REF_CAT2(a, b)
extern "C" __attribute__((visibility("default"))) void REF_CAT(ref_, VARIANT)(RefArgs *a) { harness_entry(a); }
