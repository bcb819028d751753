/* TEST INFRASTRUCTURE — oracle build shim, never part of the product library.
 *
 * Declaration-only stand-in for <gmp.h>. The image ships the GMP runtime
 * (/lib/x86_64-linux-gnu/libgmp.so.10) but no development header. The reference
 * sources under /root/reference/rust-rapidsnark/rapidsnark/src use a small set of
 * mpz_ and mpn_ entry points; every one of them is an exported `T` symbol of the
 * system library, so declaring the prototypes with the x86-64 LP64 type widths is
 * enough to compile and link the reference unmodified (oracle/Makefile).
 */
#ifndef KZP_ORACLE_GMP_SHIM_H
#define KZP_ORACLE_GMP_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned long mp_limb_t;
typedef long mp_limb_signed_t;
typedef long mp_size_t;
typedef unsigned long mp_bitcnt_t;
typedef mp_limb_t* mp_ptr;
typedef const mp_limb_t* mp_srcptr;

typedef struct
{
    int _mp_alloc;
    int _mp_size;
    mp_limb_t* _mp_d;
} __mpz_struct;

typedef __mpz_struct mpz_t[1];
typedef __mpz_struct* mpz_ptr;
typedef const __mpz_struct* mpz_srcptr;

#define GMP_LIMB_BITS 64
#define GMP_NUMB_BITS 64

/* mpz */
#define mpz_init __gmpz_init
void mpz_init(mpz_ptr);
#define mpz_clear __gmpz_clear
void mpz_clear(mpz_ptr);
#define mpz_set __gmpz_set
void mpz_set(mpz_ptr, mpz_srcptr);
#define mpz_set_ui __gmpz_set_ui
void mpz_set_ui(mpz_ptr, unsigned long);
#define mpz_set_si __gmpz_set_si
void mpz_set_si(mpz_ptr, long);
#define mpz_set_str __gmpz_set_str
int mpz_set_str(mpz_ptr, const char*, int);
#define mpz_init_set_str __gmpz_init_set_str
int mpz_init_set_str(mpz_ptr, const char*, int);
#define mpz_init_set_ui __gmpz_init_set_ui
void mpz_init_set_ui(mpz_ptr, unsigned long);
#define mpz_init_set_si __gmpz_init_set_si
void mpz_init_set_si(mpz_ptr, long);
#define mpz_get_str __gmpz_get_str
char* mpz_get_str(char*, int, mpz_srcptr);
#define mpz_get_si __gmpz_get_si
long mpz_get_si(mpz_srcptr);
#define mpz_fits_sint_p __gmpz_fits_sint_p
int mpz_fits_sint_p(mpz_srcptr);
#define mpz_cmp __gmpz_cmp
int mpz_cmp(mpz_srcptr, mpz_srcptr);
#define mpz_cmp_ui __gmpz_cmp_ui
int mpz_cmp_ui(mpz_srcptr, unsigned long);
#define mpz_cmp_si __gmpz_cmp_si
int mpz_cmp_si(mpz_srcptr, long);
#define mpz_add __gmpz_add
void mpz_add(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_add_ui __gmpz_add_ui
void mpz_add_ui(mpz_ptr, mpz_srcptr, unsigned long);
#define mpz_sub __gmpz_sub
void mpz_sub(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_sub_ui __gmpz_sub_ui
void mpz_sub_ui(mpz_ptr, mpz_srcptr, unsigned long);
#define mpz_mul __gmpz_mul
void mpz_mul(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_mod __gmpz_mod
void mpz_mod(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_mul_2exp __gmpz_mul_2exp
void mpz_mul_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
#define mpz_fdiv_q_2exp __gmpz_fdiv_q_2exp
void mpz_fdiv_q_2exp(mpz_ptr, mpz_srcptr, mp_bitcnt_t);
#define mpz_fdiv_q __gmpz_fdiv_q
void mpz_fdiv_q(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_fdiv_r __gmpz_fdiv_r
void mpz_fdiv_r(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_powm __gmpz_powm
void mpz_powm(mpz_ptr, mpz_srcptr, mpz_srcptr, mpz_srcptr);
#define mpz_invert __gmpz_invert
int mpz_invert(mpz_ptr, mpz_srcptr, mpz_srcptr);
#define mpz_tstbit __gmpz_tstbit
int mpz_tstbit(mpz_srcptr, mp_bitcnt_t);
#define mpz_sizeinbase __gmpz_sizeinbase
size_t mpz_sizeinbase(mpz_srcptr, int);
#define mpz_import __gmpz_import
void mpz_import(mpz_ptr, size_t, int, size_t, int, size_t, const void*);
#define mpz_export __gmpz_export
void* mpz_export(void*, size_t*, int, size_t, int, size_t, mpz_srcptr);

/* mpn */
#define mpn_add_n __gmpn_add_n
mp_limb_t mpn_add_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
#define mpn_sub_n __gmpn_sub_n
mp_limb_t mpn_sub_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
#define mpn_add __gmpn_add
mp_limb_t mpn_add(mp_ptr, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);
#define mpn_sub __gmpn_sub
mp_limb_t mpn_sub(mp_ptr, mp_srcptr, mp_size_t, mp_srcptr, mp_size_t);
#define mpn_add_1 __gmpn_add_1
mp_limb_t mpn_add_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
#define mpn_sub_1 __gmpn_sub_1
mp_limb_t mpn_sub_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
#define mpn_mul_1 __gmpn_mul_1
mp_limb_t mpn_mul_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
#define mpn_addmul_1 __gmpn_addmul_1
mp_limb_t mpn_addmul_1(mp_ptr, mp_srcptr, mp_size_t, mp_limb_t);
#define mpn_cmp __gmpn_cmp
int mpn_cmp(mp_srcptr, mp_srcptr, mp_size_t);
#define mpn_zero_p __gmpn_zero_p
int mpn_zero_p(mp_srcptr, mp_size_t);
#define mpn_copyi __gmpn_copyi
void mpn_copyi(mp_ptr, mp_srcptr, mp_size_t);
#define mpn_and_n __gmpn_and_n
void mpn_and_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
#define mpn_ior_n __gmpn_ior_n
void mpn_ior_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
#define mpn_xor_n __gmpn_xor_n
void mpn_xor_n(mp_ptr, mp_srcptr, mp_srcptr, mp_size_t);
#define mpn_com __gmpn_com
void mpn_com(mp_ptr, mp_srcptr, mp_size_t);
#define mpn_lshift __gmpn_lshift
mp_limb_t mpn_lshift(mp_ptr, mp_srcptr, mp_size_t, unsigned int);
#define mpn_rshift __gmpn_rshift
mp_limb_t mpn_rshift(mp_ptr, mp_srcptr, mp_size_t, unsigned int);

#ifdef __cplusplus
}
#endif

#endif
