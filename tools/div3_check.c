// gcc -O2 -mfma -ffp-contract=off -o div3_check div3_check.c -lm : x/3.0 against the FMA sequence of csrc/assembly.cu div3()
#include <stdio.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <stdlib.h>
static inline double fast3(double x){ const double r = 1.0/3.0; double q0 = x*r; double rem = fma(-3.0, q0, x); return fma(rem, r, q0); }
static uint64_t s = 88172645463325252ULL;
static inline uint64_t rnd(){ s ^= s<<13; s ^= s>>7; s ^= s<<17; return s; }
int main(){
  long bad=0; long n=0;
  for (long i=0;i<400000000L;i++){
    uint64_t b = rnd();
    // exponent in [1023-900, 1023+900]
    uint64_t e = 123 + (rnd() % 1801);
    uint64_t bits = (b & 0x800FFFFFFFFFFFFFULL) | (e<<52);
    double x; memcpy(&x,&bits,8);
    double a = x/3.0, f = fast3(x);
    if (memcmp(&a,&f,8)) { if (bad<5) printf("mismatch x=%a a=%a f=%a\n",x,a,f); bad++; }
    n++;
  }
  // structured mantissas: all-ones tails, near multiples of 3
  for (uint64_t m=0;m<(1ULL<<24);m++){
    uint64_t mant = (m<<28) | (m & 0xFFFFFFF);
    uint64_t bits = (1023ULL<<52) | (mant & 0xFFFFFFFFFFFFFULL);
    double x; memcpy(&x,&bits,8);
    for (int k=0;k<3;k++){ double y = x*(k+1); double a=y/3.0,f=fast3(y); if (memcmp(&a,&f,8)) {bad++; if (bad<5) printf("mm %a\n",y);} n++; }
  }
  for (uint64_t m=0;m<(1ULL<<22);m++){ double x = (double)(3*m) ; double a=x/3.0,f=fast3(x); if (memcmp(&a,&f,8)) bad++; x=(double)(3*m+1)*0x1p30+ (double)m; a=x/3.0; f=fast3(x); if (memcmp(&a,&f,8)) bad++; n+=2; }
  printf("tested %ld, mismatches %ld\n", n, bad);
  return 0;
}
