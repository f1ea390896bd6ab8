#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <omp.h>
int main(){
    const double PI=3.14159265358979323846; const double RCP=1.0/PI;
    long long bad=0, badd=0, n=0;
    #pragma omp parallel for reduction(+:bad,badd,n) schedule(static)
    for (uint64_t bits=0; bits<=0x40800000ull; bits++){ // 0 .. 4.0f
        uint32_t b=(uint32_t)bits; float c; memcpy(&c,&b,4);
        double cd=c;
        double ref=cd/PI;
        double q0=cd*RCP;
        double r=fma(-q0,PI,cd);
        double q1=fma(r,RCP,q0);
        if (q1!=ref) badd++;
        if ((float)q1!=(float)ref) bad++;
        n++;
    }
    printf("checked %lld floats in [0,4]: double mismatches %lld, float-result mismatches %lld\n", n, badd, bad);
    // also rcp equivalence: (float)(1.0/(double)x) == 1.0f/x for all normal floats with normal reciprocal
    long long badr=0, nr=0;
    #pragma omp parallel for reduction(+:badr,nr) schedule(static)
    for (uint64_t bits=0x00800000ull; bits<0x7f000000ull; bits++){
        uint32_t b=(uint32_t)bits; float x; memcpy(&x,&b,4);
        float a=(float)(1.0/(double)x); float c=1.0f/x;
        if (a!=c) badr++;
        nr++;
    }
    printf("rcp: checked %lld positive normal floats, mismatches %lld\n", nr, badr);
    return 0;
}
