// Host emulation, operation by operation, of the device Gaussian pk_gauss<double> (kamphir_b200/csrc/phylok_dev.cu):
// lambda * 2^(-E/64) from a 64-entry table with biased high words, a degree-4 polynomial and an integer exponent add.
// Prints the maximum relative error against expl() over 4e6 arguments per decayFactor; tests/test_gauss_host.py builds
// it with the coefficients taken from the .cu file (-DPK_C1=... -DPK_C4=...) and asserts the bound quoted in DESIGN.md.
//   gcc -O2 -ffp-contract=off -o gauss_check gauss_check.c -lm && ./gauss_check
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static double tab_lo_hi[64]; static int32_t tabhi[64]; static uint32_t tablo[64];
#ifndef PK_C1
#define PK_C1 -0x1.62e42fefa2417p-7
#define PK_C2 0x1.ebfbdff82bb76p-15
#define PK_C3 -0x1.c6b0b9214b9aep-23
#define PK_C4 0x1.3b2acb2c36b23p-31
#endif
static const double C1=PK_C1, C2=PK_C2, C3=PK_C3, C4=PK_C4;
static inline int32_t hi(double x){uint64_t u;memcpy(&u,&x,8);return (int32_t)(u>>32);}
static inline uint32_t lo(double x){uint64_t u;memcpy(&u,&x,8);return (uint32_t)u;}
static inline double mk(int32_t h,uint32_t l){uint64_t u=((uint64_t)(uint32_t)h<<32)|l;double x;memcpy(&x,&u,8);return x;}
double gauss(double E, int32_t clamp_hi){
  const double MAGIC=6755399441055744.0;
  uint32_t eh=(uint32_t)hi(E); if(eh>(uint32_t)clamp_hi) eh=clamp_hi;
  double Ec=mk((int32_t)eh,lo(E));
  double t=MAGIC-Ec; int32_t k=(int32_t)lo(t); double tn=t-MAGIC; double r=tn+Ec;
  double p=fma(C4,r,C3); p=fma(p,r,C2); p=fma(p,r,C1); p=fma(p,r,1.0);
  int j=k&63; int32_t h=tabhi[j]+(int32_t)((uint32_t)k<<14);
  return p*mk(h,tablo[j]);
}
int main(){
  int bad=0;
  double lambdas[]={0.2,0.1,0.5,1.0,3.7,1e-5,1e-30};
  for(int li=0;li<7;li++){
    double lambda=lambdas[li];
    for(int j=0;j<64;j++){double T=lambda*exp2(j/64.0); tabhi[j]=hi(T)-(j<<14); tablo[j]=lo(T);}
    int fl=(int)floor(log2(lambda)); int emin=2-1023-fl; double Emax=-64.0*emin-1; int32_t clamp_hi=hi(Emax)-1;
    double worst=0; double worstE=0;
    srand(1);
    for(int i=0;i<4000000;i++){
      double u=rand()/(double)RAND_MAX; double E;
      if(i%4==0) E=u*70; else if(i%4==1) E=u*3000; else if (i%4==2) E=u*u*40000; else E=u*1e-3;
      double got=gauss(E,clamp_hi); long double want=(long double)lambda*expl(-(long double)E*logl(2.0L)/64.0L);
      double rel=fabs((double)((got-want)/want)); if(rel>worst){worst=rel;worstE=E;}
    }
    if (worst > 1e-14) bad = 1;
    printf("lambda %g: max rel err %.3g at E=%g ; E=0 -> %.17g ; clamp Emax=%g, huge: %g %g %g\n",lambda,worst,worstE,gauss(0,clamp_hi),Emax,gauss(Emax,clamp_hi),gauss(1e9,clamp_hi),gauss(1e300,clamp_hi));
  }
  return bad;
}
