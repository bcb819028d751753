/* TEST INFRASTRUCTURE — oracle build shim, never part of the product library.
 *
 * The reference samples the Groth16 blinding scalars r,s through
 * randombytes_buf() (random_generator.hpp:4-8, call sites groth16.cpp:302,311).
 * Compiling groth16.cpp with -DUSE_SODIUM makes it include <sodium.h>; this file
 * answers that include with a replayable source so parity tests can fix (r,s)
 * WITHOUT modifying any reference source:
 *   - if kzp_oracle_fixed_rs_set(r32, s32) was called, successive calls hand out
 *     r then s (32 bytes each, little-endian canonical values < the Fr modulus
 *     with the top two bits clear so the reference's rejection loop accepts);
 *   - otherwise bytes come from std::random_device like the reference default.
 */
#ifndef KZP_ORACLE_SODIUM_SHIM_H
#define KZP_ORACLE_SODIUM_SHIM_H

#include <cstddef>
#include <cstdint>

extern "C" void kzp_oracle_fixed_rs_set(const uint8_t* r32, const uint8_t* s32);
extern "C" void kzp_oracle_fixed_rs_clear(void);
extern "C" void randombytes_buf(void* const buf, const size_t size);

#endif
