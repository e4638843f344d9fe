/*
 * CPU model of the tile kernel's sorted-order fast path (aloception_oss_b200/csrc/sortv_capi.cu, phase B), test infrastructure:
 * the same checks (irreflexive, every pair ordered exactly one way, in-degrees a permutation, every candidate before the start
 * value, one round per candidate) and the same joint evaluation of before(a, b) / before(b, a) in host float arithmetic
 * (-ffp-contract=off, the GPU's one fma written out).  tests/test_sortv_oracle.py runs it against the scalar oracle
 * (oracle/sortv_oracle.c): whenever the fast path applies its order must be the order of the reference's selection rounds.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#define EPS 1e-8
static const float kF = (float)EPS;
static int above(void){ return (double)kF > EPS; }
static int lt_eps(float f){ return above() ? (f < kF) : (f <= kF); }
static int gt_eps(float f){ return above() ? (f >= kF) : (f > kF); }
static float q_of(float x, float y){ float n = (float)((double)fmaf(x,x,y*y) + EPS); return (fabsf(x)*x)/n; }
static int before_q(float x1,float y1,float q1,float x2,float y2,float q2){
  int tie = lt_eps(fabsf(x1-x2)) & lt_eps(fabsf(y2-y1));
  int p1=y1>0.f,n1=y1<0.f,p2=y2>0.f,n2=y2<0.f; float d=q1-q2;
  return !tie & ((p1&n2)|(p1&p2&gt_eps(d))|(n1&n2&lt_eps(d)));
}
/* returns 1 if fast path applies; writes order */
int fast_order(const float* v, const uint8_t* mk, int nv_in, int m, int* out, int* applied){
  float cx[32],cy[32],cq[32]; int ck[32]; int c=0;
  for(int k=0;k<m;k++) if(mk[k]){cx[c]=v[2*k];cy[c]=v[2*k+1];cq[c]=q_of(cx[c],cy[c]);ck[c]=k;c++;}
  int nv = nv_in>8?8:nv_in;
  *applied=0;
  if(nv_in<3||c>8) return 0;
  float y0=-kF,q0=q_of(1.f,y0);
  int total = nv==c; unsigned rank[8]={0}; 
  for(int a=0;a<c;a++){ total &= before_q(cx[a],cy[a],cq[a],1.f,y0,q0) & !before_q(cx[a],cy[a],cq[a],cx[a],cy[a],cq[a]);
    for(int b=a+1;b<c;b++){ int tie=lt_eps(fabsf(cx[a]-cx[b]))&lt_eps(fabsf(cy[b]-cy[a]));
      int pa=cy[a]>0.f,na=cy[a]<0.f,pb=cy[b]>0.f,nb=cy[b]<0.f; float d=cq[a]-cq[b],nd=-d;
      int ab=!tie&((pa&nb)|(pa&pb&gt_eps(d))|(na&nb&lt_eps(d)));
      int ba=!tie&((pb&na)|(pb&pa&gt_eps(nd))|(nb&na&lt_eps(nd)));
      int ab2=before_q(cx[a],cy[a],cq[a],cx[b],cy[b],cq[b]), ba2=before_q(cx[b],cy[b],cq[b],cx[a],cy[a],cq[a]);
      if(ab!=ab2||ba!=ba2) printf("JOINT MISMATCH ab %d %d ba %d %d  (%g,%g,%g) (%g,%g,%g)\n",ab,ab2,ba,ba2,cx[a],cy[a],cq[a],cx[b],cy[b],cq[b]);
      total &= ab!=ba; rank[b]+=ab; rank[a]+=ba; } }
  unsigned seen=0; for(int a=0;a<c;a++){ seen|=1u<<rank[a]; }
  total &= seen==(1u<<c)-1u;
  if(!total) return 0;
  for(int a=0;a<c;a++) out[rank[a]]=ck[a];
  *applied=1; return 1;
}
