// sq_glibc_data.h -- coefficient tables of glibc 2.39's float sincosf / powf (x86-64), as 64-bit patterns of doubles.
//
// The reference's sampler (fast_sampler/sampling.cpp:59-67) calls libm's cosf, sinf and powf; to reproduce its
// discrete decisions bit for bit the device evaluates the SAME double-precision algorithms (sq_math.cuh) with the
// SAME tables.  The numbers below were read out of /usr/lib/x86_64-linux-gnu/libm.so.6 (glibc 2.39-0ubuntu8.5) of
// the build image (objdump of __cosf_fma / __sinf_fma / __powf_fma for the addresses, then the .rodata bytes); they
// are the published tables of ARM's optimized-routines sincosf/powf (log2 table 2^4 entries, exp2 table 2^5 entries)
// that glibc adopted in 2.28.  tests/test_sq_math.py checks the host build of sq_math.cuh against libm itself.
#pragma once
#include <stdint.h>

// sincos_t x2: sign[4], hpi_inv, hpi, then c0, c1, s1, c2, s2, c3, s3, c4 (the order the compiled code indexes them)
#define SQ_GLIBC_SINCOS { \
    0x3ff0000000000000ull, 0xbff0000000000000ull, 0xbff0000000000000ull, 0x3ff0000000000000ull, \
    0x41645f306dc9c883ull, 0x3ff921fb54442d18ull, 0x3ff0000000000000ull, 0xbfdffffffd0c621cull, \
    0xbfc555545995a603ull, 0x3fa55553e1068f19ull, 0x3f81107605230bc4ull, 0xbf56c087e89a359dull, \
    0xbf2994eb3774cf24ull, 0x3ef99343027bf8c3ull, 0x3ff0000000000000ull, 0xbff0000000000000ull, \
    0xbff0000000000000ull, 0x3ff0000000000000ull, 0x41645f306dc9c883ull, 0x3ff921fb54442d18ull, \
    0xbff0000000000000ull, 0x3fdffffffd0c621cull, 0xbfc555545995a603ull, 0xbfa55553e1068f19ull, \
    0x3f81107605230bc4ull, 0x3f56c087e89a359dull, 0xbf2994eb3774cf24ull, 0xbef99343027bf8c3ull }
// powf: log2 table {invc, logc} x16
#define SQ_GLIBC_LOG2TAB { \
    0x3ff661ec79f8f3beull, 0xbfdefec65b963019ull, 0x3ff571ed4aaf883dull, 0xbfdb0b6832d4fca4ull, \
    0x3ff49539f0f010b0ull, 0xbfd7418b0a1fb77bull, 0x3ff3c995b0b80385ull, 0xbfd39de91a6dcf7bull, \
    0x3ff30d190c8864a5ull, 0xbfd01d9bf3f2b631ull, 0x3ff25e227b0b8ea0ull, 0xbfc97c1d1b3b7af0ull, \
    0x3ff1bb4a4a1a343full, 0xbfc2f9e393af3c9full, 0x3ff12358f08ae5baull, 0xbfb960cbbf788d5cull, \
    0x3ff0953f419900a7ull, 0xbfaa6f9db6475fceull, 0x3ff0000000000000ull, 0x0000000000000000ull, \
    0x3fee608cfd9a47acull, 0x3fb338ca9f24f53dull, 0x3feca4b31f026aa0ull, 0x3fc476a9543891baull, \
    0x3feb2036576afce6ull, 0x3fce840b4ac4e4d2ull, 0x3fe9c2d163a1aa2dull, 0x3fd40645f0c6651cull, \
    0x3fe886e6037841edull, 0x3fd88e9c2c1b9ff8ull, 0x3fe767dcf5534862ull, 0x3fdce0a44eb17bccull }
// powf: log2 polynomial A[0..4], then exp2f SHIFT and polynomial C[0..2]
#define SQ_GLIBC_POLY { \
    0x3fd27616c9496e0bull, 0xbfd71969a075c67aull, 0x3fdec70a6ca7baddull, 0xbfe7154748bef6c8ull, \
    0x3ff71547652ab82bull, 0x42e8000000000000ull, 0x3fac6af84b912394ull, 0x3fcebfce50fac4f3ull, \
    0x3fe62e42ff0c52d6ull }
// exp2f table x32
#define SQ_GLIBC_EXP2TAB { \
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, \
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, \
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull, \
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull, \
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull, \
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, \
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, \
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull }
