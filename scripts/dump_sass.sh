#!/bin/bash
# Dump SASS listings (instruction text, no encodings) of the hot kernels from the built library.
cd "$(dirname "$0")/.."
LIB=slam-1_b200/libslammatch.so
dump() {  # $1 = substring of the mangled name, $2 = output file
  cuobjdump -sass $LIB | awk -v pat="$1" '/Function : /{f = index($0, pat) > 0} f{print}' \
    | grep -E "Function :|^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##; s/[ \t]+$//' > profiles/sass/$2
  echo "$2: $(wc -l < profiles/sass/$2) lines"
}
if [ "$1" != "r2" ]; then   # "dump_sass.sh r2" leaves the round-1 listings as they were committed
dump "knn2_tc2_kernelILi4ELb0" r1_knn2_tc2_kernel_mt4.sass
dump "knn2_tc_kernelILi1" r1_knn2_tc_kernel_mt1.sass
dump "knn2_popc_kernel" r1_knn2_popc_kernel.sass
dump "knn2_bmma_kernel" r1_knn2_bmma_kernel.sass
dump "knn2_stream_kernelILi1" r1_knn2_stream_kernel_nq1.sass
dump "tc_refine_kernelILi32" r1_tc_refine_kernel_g32.sass
dump "knn2_tc2_kernelILi4ELb1" r1_knn2_tc2_kernel_mt4_chain.sass
dump "tc_refine_frame_kernel" r1_tc_refine_frame_kernel.sass
dump "knn2_frame_kernelILi2ELi8" r1_knn2_frame_kernel_kq2_w8.sass
dump "vocab_majority_kernel" r1_vocab_majority_kernel.sass
grep -c "UTCQMMA" profiles/sass/r1_knn2_tc2_kernel_mt4.sass
fi
# round 2
dump "knn2_tc4_kernelILi120ELb0ELb0" r2_knn2_tc4_kernel_ch120.sass
dump "knn2_tc4_kernelILi40ELb0ELb0" r2_knn2_tc4_kernel_ch40.sass
dump "exchange_wait_merge_kernel" r2_exchange_wait_merge_kernel.sass
dump "tc_refine_kernelILi32" r2_tc_refine_kernel_g32.sass
echo "UTCOMMA $(grep -c UTCOMMA profiles/sass/r2_knn2_tc4_kernel_ch120.sass)  UBLKCP $(grep -c UBLKCP profiles/sass/r2_knn2_tc4_kernel_ch120.sass)  LDTM $(grep -c LDTM profiles/sass/r2_knn2_tc4_kernel_ch120.sass)  LDL $(grep -c 'LDL' profiles/sass/r2_knn2_tc4_kernel_ch120.sass)"
# round 2, closing state: the reworked refine kernels (carry-save POPC), the two-phase exchange kernels, the masked search,
# the wide chi-square scan, the chained mxf4 instantiation of config 3
dump "tc_refine_kernelILi8" r2_tc_refine_kernel_g8.sass
dump "tc_refine_frame_kernel" r2_tc_refine_frame_kernel.sass
dump "tc_refine_owned_kernel" r2_tc_refine_owned_kernel.sass
dump "tc_chunk_keys_flat_kernel" r2_tc_chunk_keys_flat_kernel.sass
dump "knn2_masked_kernel" r2_knn2_masked_kernel.sass
dump "chi2_scan_wide_kernel" r2_chi2_scan_wide_kernel.sass
dump "knn2_tc4_kernelILi20" r2_knn2_tc4_kernel_ch20_chain.sass
echo "refine g32: POPC $(grep -c POPC profiles/sass/r2_tc_refine_kernel_g32.sass)  masked: POPC $(grep -c POPC profiles/sass/r2_knn2_masked_kernel.sass) LDG $(grep -c LDG profiles/sass/r2_knn2_masked_kernel.sass)"
