// Test program (tests/ only): the three-operation division step used by the bundle-merge
// kernel (voxblox_b200/csrc/vbx_tsdf.cu exact_div) against IEEE-754 division on random and
// adversarial operands (all-ones / near-power-of-two mantissas, exponents -60..60).
// Build: gcc -O2 -mfma -ffp-contract=off.  Prints "n=... bad=0".
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static inline float asf(uint32_t u){float f; memcpy(&f,&u,4); return f;}
static inline uint32_t asu(float f){uint32_t u; memcpy(&u,&f,4); return u;}
static uint64_t s=88172645463325252ull;
static inline uint64_t rnd(){ s^=s<<13; s^=s>>7; s^=s<<17; return s;}
int main(){
  long bad=0, n=0, badones=0;
  for(long it=0; it<100000000L; ++it){
    uint64_t r=rnd();
    // a: random mantissa, exponent in [-20,20]; b likewise, positive
    uint32_t ma=(uint32_t)(r&0x7fffff), mb=(uint32_t)((r>>23)&0x7fffff);
    int ea=(int)((r>>46)%121)-60, eb=(int)((r>>52)%121)-60;
    uint32_t sa=(uint32_t)((r>>60)&1);
    if((it&7)==0) mb = 0x7fffff & ~(uint32_t)(rnd()&0x3); // near all-ones mantissas
    if((it&7)==1) mb = (uint32_t)(rnd()&0x7);              // near power of two
    if((it&15)==2) ma = 0x7fffff & ~(uint32_t)(rnd()&0x3);
    float a=asf((sa<<31)|((uint32_t)(ea+127)<<23)|ma), b=asf(((uint32_t)(eb+127)<<23)|mb);
    float ref=a/b;
    float y=1.0f/b;              // correctly rounded reciprocal (== __frcp_rn)
    float q=a*y;
    float rem=fmaf(-q,b,a);
    float q2=fmaf(rem,y,q);
    ++n;
    if(asu(q2)!=asu(ref)){ ++bad; if(mb==0x7fffff) ++badones; if(bad<10) printf("mismatch a=%a b=%a ref=%a got=%a mb=%x\n",a,b,ref,q2,mb);}
  }
  printf("n=%ld bad=%ld (all-ones mantissa: %ld)\n",n,bad,badones);
  return 0;
}
