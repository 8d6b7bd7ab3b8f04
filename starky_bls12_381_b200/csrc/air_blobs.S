/* The five standard constraint programs (tools/airgen output, committed as ../air/<name>.airbin.xz and unpacked by the
 * Makefile into ../air/_unpacked/, which is on the assembler's include path) linked into libstarkyb200.so, so that
 * sb_prove needs no file at run time (quotient.cu: air_get).  They replace the per-point callback into
 * Stark::eval_packed_generic (/root/reference/src/fp12_mul.rs:58, calc_pairing_precomp.rs:376, miller_loop.rs:644,
 * final_exponentiate.rs:907, ecc_aggregate.rs:92). */
    .section .rodata
    .global sb_airbin_fp12_mul
    .global sb_airbin_fp12_mul_end
    .balign 64
sb_airbin_fp12_mul:
    .incbin "fp12_mul.airbin"
sb_airbin_fp12_mul_end:
    .global sb_airbin_pairing_precomp
    .global sb_airbin_pairing_precomp_end
    .balign 64
sb_airbin_pairing_precomp:
    .incbin "pairing_precomp.airbin"
sb_airbin_pairing_precomp_end:
    .global sb_airbin_miller_loop
    .global sb_airbin_miller_loop_end
    .balign 64
sb_airbin_miller_loop:
    .incbin "miller_loop.airbin"
sb_airbin_miller_loop_end:
    .global sb_airbin_final_exp
    .global sb_airbin_final_exp_end
    .balign 64
sb_airbin_final_exp:
    .incbin "final_exp.airbin"
sb_airbin_final_exp_end:
    .global sb_airbin_ecc_agg
    .global sb_airbin_ecc_agg_end
    .balign 64
sb_airbin_ecc_agg:
    .incbin "ecc_agg.airbin"
sb_airbin_ecc_agg_end:
    .section .note.GNU-stack,"",@progbits
