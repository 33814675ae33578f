#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <random>
extern "C" int tfbs_debug_hits_plan_checksum(const double*, const int32_t*, int32_t, int32_t, const float*, uint64_t*, int32_t*);
extern "C" int tfbs_debug_hits_filter(const double*, const int32_t*, int32_t, int32_t, const float*, int32_t, const uint8_t*, int64_t, uint8_t*, int32_t*, double*, int32_t*);
extern "C" int tfbs_debug_dense_tables(const double*, int32_t, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*, int64_t);
extern "C" int tfbs_debug_swar4(uint32_t, uint32_t*, uint32_t*);
extern "C" int tfbs_debug_direct_lookup(const double*, const int32_t*, int32_t, int32_t, const float*, const uint8_t*, int64_t, uint8_t*, int64_t*);
int main(){
  std::mt19937 rng(7);
  for(int cs=0; cs<14; cs++){
    const int M = (int[]){1,31,32,33,64,65,97,130,200,333,800,40,96,513}[cs];
    const int lmin = (int[]){1,1,6,6,6,1,10,15,6,6,6,25,4,6}[cs], lmax=(int[]){1,32,30,30,14,12,32,30,30,24,30,32,14,30}[cs];
    const int Lmax=lmax;
    std::vector<double> lo((size_t)M*Lmax*4, 0.0); std::vector<int32_t> lens(M); std::vector<float> thr(M);
    std::uniform_real_distribution<double> U(0.02,1.0);
    const double frac = (double[]){0.5,0.6,0.7,0.8,0.9}[cs%5];
    for(int m=0;m<M;m++){ lens[m]=lmin+rng()%(lmax-lmin+1); double best=0; for(int j=0;j<lens[m];j++){ double p[4],s=0; for(int b=0;b<4;b++){p[b]=std::pow(U(rng),4.0); s+=p[b];} double mx=-1e9; for(int b=0;b<4;b++){ double v=std::log2(p[b]/s/0.25); if(rng()%97==0) v=-INFINITY; lo[((size_t)m*Lmax+j)*4+b]=v; if(std::isfinite(v)) mx=std::max(mx,v);} best+=mx;} thr[m]=(float)(best*frac); if(m%17==3) thr[m]=-INFINITY; if(m%19==4) thr[m]=INFINITY; if(m%23==5) thr[m]=NAN; }
    uint64_t out[6]; std::vector<int32_t> plan((size_t)M*4);
    int rc=tfbs_debug_hits_plan_checksum(lo.data(),lens.data(),M,Lmax,thr.data(),out,plan.data());
    const int64_t n=700; std::vector<uint8_t> codes(n); for(auto&c:codes) c=rng()&3;
    std::vector<uint8_t> cand((size_t)n*M*2); std::vector<double> exp(M); int32_t nb=0; int64_t info[4];
    int rc2=0; for(int form=0; form<=4; form++) rc2|=tfbs_debug_hits_filter(lo.data(),lens.data(),M,Lmax,thr.data(),form,codes.data(),n,cand.data(),plan.data(),exp.data(),&nb);
    int rc3=tfbs_debug_direct_lookup(lo.data(),lens.data(),M,Lmax,thr.data(),codes.data(),n,cand.data(),info);
    // dense-scan fixed-point tables of every motif (tfbs_dense_quantise, run at library upload)
    int rc4 = 0;
    std::vector<int32_t> ent((size_t)11 * 256 * 2);
    for (int m = 0; m < M; m += (M > 100 ? 7 : 1)) {
      std::vector<double> one((size_t)lens[m] * 4);
      for (int j = 0; j < lens[m]; j++) for (int b = 0; b < 4; b++) one[(size_t)j * 4 + b] = lo[((size_t)m * Lmax + j) * 4 + b];
      int32_t K, G, e, neg;
      rc4 |= tfbs_debug_dense_tables(one.data(), lens[m], &K, &G, &e, &neg, ent.data(), (int64_t)ent.size());
    }
    uint32_t c8, v4; rc4 |= tfbs_debug_swar4(0x4E544341u ^ (uint32_t)rng(), &c8, &v4);
    rc2 |= rc4;
    printf("case %d M %d L %d-%d: rc %d %d %d blocks %llu direct %llu keys %llu\n",cs,M,lmin,lmax,rc,rc2,rc3,(unsigned long long)out[0],(unsigned long long)out[4],(unsigned long long)out[5]);
  }
  return 0; }
