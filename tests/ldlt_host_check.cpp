#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "linalg.cuh"
extern "C" void oracle_ldlt_solve(int n, const float* A, const float* rhs, float* x);
template<int N> int run(int trials){
  int bad=0;
  for(int t=0;t<trials;t++){
    float A[N*N], B[N*N], rhs[N], x1[N], x2[N];
    float G[N*N];
    for(int i=0;i<N*N;i++) G[i]=(float)rand()/RAND_MAX*2-1;
    int mode=t%4;
    for(int i=0;i<N;i++)for(int j=0;j<N;j++){ float s=0; for(int k=0;k<N;k++) s+=G[k*N+i]*G[k*N+j]; A[j*N+i]=s; }
    if(mode==1) for(int i=0;i<N;i++) A[i*N+i]-=1.0f;           // indefinite
    if(mode==2){ for(int i=0;i<N;i++){A[i*N+0]=0;A[0*N+i]=0;} } // singular
    if(mode==3){ for(int i=0;i<N;i++)for(int j=0;j<N;j++) A[j*N+i]*= (1+100*i)*(1+100*j);} // badly scaled
    for(int i=0;i<N;i++) rhs[i]=(float)rand()/RAND_MAX*2-1;
    memcpy(B,A,sizeof(A));
    oracle_ldlt_solve(N,A,rhs,x1);
    vo::ldlt_solve_dev<N>(B,rhs,x2);
    if(memcmp(x1,x2,sizeof(x1))!=0){ bad++; if(bad<3){printf("N=%d mode=%d mismatch:",N,mode); for(int i=0;i<N;i++)printf(" %g/%g",x1[i],x2[i]); printf("\n");} }
  }
  return bad;
}
int main(){ int b6=run<6>(20000), b2=run<2>(20000); printf("bad6=%d bad2=%d\n",b6,b2); return b6+b2?1:0; }
