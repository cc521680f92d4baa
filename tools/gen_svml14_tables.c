// Generate and verify compact tables for vrsqrt14ps / vrcp14ps (AVX-512F): both depend only on the top
// mantissa bits of the input (plus the exponent parity for rsqrt) except for exact powers, which are exact.
//   gcc -O2 -mavx512f -o gen tools/gen_svml14_tables.c && ./gen tables.bin   (then see csrc/svml14_tables.inc)
#include <immintrin.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static float rsqrt14(float x){ __m512 v=_mm512_set1_ps(x); v=_mm512_rsqrt14_ps(v); float o[16]; _mm512_storeu_ps(o,v); return o[0]; }
static float rcp14(float x){ __m512 v=_mm512_set1_ps(x); v=_mm512_rcp14_ps(v); float o[16]; _mm512_storeu_ps(o,v); return o[0]; }
static uint32_t fb(float f){uint32_t u; memcpy(&u,&f,4); return u;}
static float bf(uint32_t u){float f; memcpy(&f,&u,4); return f;}
int main(int argc,char**argv){
  // rsqrt table: index = (exponent parity << 15) | top 15 mantissa bits; value = mantissa+exponent of result for x in [1,4)
  static uint32_t rsq[1<<16], rcp[1<<16];
  for (uint32_t par=0; par<2; ++par) for (uint32_t t=0;t<(1u<<15);++t){
    uint32_t b=((127u+par)<<23)|(t<<8)|0x80; rsq[(par<<15)|t]=fb(rsqrt14(bf(b))); }
  for (uint32_t t=0;t<(1u<<16);++t){ uint32_t b=(127u<<23)|(t<<7)|0x40; rcp[t]=fb(rcp14(bf(b))); }
  // verify on all inputs in [1,4) / [1,2) and on other exponents (scaling)
  long bad=0;
  for (uint32_t e=127;e<=128;++e) for(uint32_t m=0;m<(1u<<23);++m){
    uint32_t b=(e<<23)|m; float want=rsqrt14(bf(b)); float got;
    if (m==0 && e==127) got=1.0f; else got=bf(rsq[((e-127)<<15)|(m>>8)]);
    if (fb(want)!=fb(got)) { if(bad<5) printf("rsq mismatch %08x want %08x got %08x\n",b,fb(want),fb(got)); bad++; }
  }
  printf("rsqrt14 table mismatches in [1,4): %ld\n",bad);
  bad=0;
  for(uint32_t m=0;m<(1u<<23);++m){ uint32_t b=(127u<<23)|m; float want=rcp14(bf(b)); float got;
    if (m==0) got=1.0f; else got=bf(rcp[m>>7]);
    if (fb(want)!=fb(got)) { if(bad<5) printf("rcp mismatch %08x want %08x got %08x\n",b,fb(want),fb(got)); bad++; } }
  printf("rcp14 table mismatches in [1,2): %ld\n",bad);
  // scaling check on random exponents: rsqrt14(x * 4^k) == rsqrt14(x) * 2^-k ; rcp14(x*2^k) == rcp14(x)*2^-k
  srand(1); bad=0;
  for (int i=0;i<20000000;++i){
    uint32_t m=((uint32_t)rand()<<8 ^ rand()) & 0x7fffff; int e=rand()%200+20; // exponent field 20..219
    uint32_t b=((uint32_t)e<<23)|m; float x=bf(b);
    // rsqrt via table
    int ue=e-127; int par=ue&1; int k=(ue-par)/2; // x = y*4^k, y in [1,4)
    float r; if (m==0 && par==0) r=1.0f; else r=bf(rsq[(par<<15)|(m>>8)]);
    uint32_t rb=fb(r)-((uint32_t)k<<23); if (fb(rsqrt14(x))!=rb) {if(bad<5)printf("rsq scale mismatch %08x\n",b); bad++;}
    float c; if(m==0) c=1.0f; else c=bf(rcp[m>>7]);
    uint32_t cb=fb(c)-((uint32_t)ue<<23); if (fb(rcp14(x))!=cb) {if(bad<5)printf("rcp scale mismatch %08x want %08x got %08x\n",b,fb(rcp14(x)),cb); bad++;}
  }
  printf("scaling mismatches: %ld\n",bad);
  if (argc>1){ FILE*f=fopen(argv[1],"wb"); fwrite(rsq,4,1<<16,f); fwrite(rcp,4,1<<16,f); fclose(f); printf("wrote %s\n",argv[1]); }
  return 0;
}
