#include <cstdio>
#include <cstdint>
#include <cmath>
#include <vector>
#include <random>
extern "C" int tfbs_debug_hits_plan_checksum(const double*, const int32_t*, int32_t, int32_t, const float*, uint64_t*, int32_t*);
extern "C" int tfbs_debug_direct_motifs(const double*, const int32_t*, int32_t, int32_t, const float*, uint8_t*);
int main(){
  const int M=200, Lmax=30;
  std::mt19937 rng(5);
  std::vector<double> lo((size_t)M*Lmax*4, 0.0); std::vector<int32_t> lens(M); std::vector<float> thr(M);
  std::uniform_real_distribution<double> U(0.02,1.0);
  for(int m=0;m<M;m++){ lens[m]=6+rng()%25; double best=0; for(int j=0;j<lens[m];j++){ double p[4],s=0; for(int b=0;b<4;b++){p[b]=std::pow(U(rng),4.0); s+=p[b];} double mx=-1e9; for(int b=0;b<4;b++){ double v=std::log2(p[b]/s/0.25); lo[((size_t)m*Lmax+j)*4+b]=v; mx=std::max(mx,v);} best+=mx;} thr[m]=(float)(best*0.55); if(m%17==3) thr[m]=-INFINITY; }
  uint64_t out[6]; std::vector<int32_t> plan(M*4);
  for(int r=0;r<3;r++){ int rc=tfbs_debug_hits_plan_checksum(lo.data(),lens.data(),M,Lmax,thr.data(),out,plan.data()); printf("rc %d blocks %llu direct %llu keys %llu tables %llx\n",rc,(unsigned long long)out[0],(unsigned long long)out[4],(unsigned long long)out[5],(unsigned long long)out[2]); }
  return 0; }
