// tr_fused_mn.h — table of k_fused_mn instantiations shared by tr_api.cu and tr_fusedmn.cu.
#pragma once
#include "tr_fused_mn.cuh"

struct TrmEntry {
    int IKC, RKS;
    const void* kern;       // void (*)(FusedMnArgs<T>)
    const void* f3_symbol;  // the translation unit's constant buffer trm_c_f3 (cudaGetSymbolAddress / cudaMemcpyToSymbol)
};

const TrmEntry* trm_entries_f32_0(int* n);
const TrmEntry* trm_entries_f32_1(int* n);
const TrmEntry* trm_entries_f32_2(int* n);
const TrmEntry* trm_entries_f64_0(int* n);
