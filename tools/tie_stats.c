// Developer statistics (not product, not oracle): walks the easy-path pyramid of one label map and counts,
// per level, step distances, mirror-tie events and the ranges of (pref, offset) at those events.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static int exact_pref(long p0,long p1){ if(!p0||!p1) return 1; long a=labs(p0),b=labs(p1); return a==b && (a&(a-1))==0; }
int main(int argc,char**argv){
  int H=512,W=512,L=16; int N=H*W;
  int32_t*lab=malloc(4*N); FILE*f=fopen(argv[1],"rb"); if(fread(lab,4,N,f)!=(size_t)N) return 1; fclose(f);
  // regions by first appearance
  int32_t*rid=malloc(4*N); int R=0; { int cap=1<<21; int32_t*k=malloc(4*cap),*v=malloc(4*cap); memset(v,0xff,4*cap);
    for(int p=0;p<N;p++){ uint32_t h=((uint32_t)lab[p]*2654435761u)&(cap-1); while(v[h]!=-1&&k[h]!=lab[p]) h=(h+1)&(cap-1); if(v[h]==-1){k[h]=lab[p];v[h]=R++;} rid[p]=v[h]; } }
  // per-region pixel lists (row-major)
  int*cnt=calloc(R+1,4); for(int p=0;p<N;p++) cnt[rid[p]+1]++; for(int r=0;r<R;r++) cnt[r+1]+=cnt[r];
  int32_t*pix=malloc(4*N); int*fill=calloc(R,4); for(int p=0;p<N;p++){ int r=rid[p]; pix[cnt[r]+fill[r]++]=p; }
  int*off=malloc(4*(R+1)); memcpy(off,cnt,4*(R+1));
  int32_t*owner=malloc(4*N); memset(owner,0xff,4*N);
  int32_t*cur=pix; int n_l=N;
  long tot_steps=0,tot_tie=0,tot_tie_nonexact=0, tot_inrange[4]={0,0,0,0};
  for(int lev=1;lev<=L;lev++){
    long steps=0, dist[6]={0}, tie=0, tie_ne=0, prefin5=0, found5=0, found5_pref5=0, tie_ne_r[4]={0}, tie_opp=0, tie_sameset=0, tie_res[3]={0};
    int32_t*next=malloc(4*(n_l/2+1)); int*noff=malloc(4*(R+1)); int32_t*path=malloc(4*n_l);
    for(int r=0;r<R;r++){
      int a=off[r], n=off[r+1]-off[r]; if(n==0) continue; int32_t*px=cur+a;
      int start=0,rmin=H,rmax=-1,cmin=W,cmax=-1;
      for(int i=0;i<n;i++){ owner[px[i]]=r; if(px[i]<px[start]) start=i; int rr=px[i]/W,cc=px[i]%W; if(rr<rmin)rmin=rr; if(rr>rmax)rmax=rr; if(cc<cmin)cmin=cc; if(cc>cmax)cmax=cc; }
      int ci=px[start]/W,cj=px[start]%W; owner[px[start]]=-1; path[a]=px[start]; long p0=0,p1=1;
      for(int t=1;t<n;t++){
        int found=0; long bdi=0,bdj=0,bd2=0,bdot=0; long adi=0,adj=0; int alt=0; int rad;
        for(rad=1;!found;rad<<=1){
          int i0=ci-rad<rmin?rmin:ci-rad,i1=ci+rad>rmax?rmax:ci+rad,j0=cj-rad<cmin?cmin:cj-rad,j1=cj+rad>cmax?cmax:cj+rad;
          for(int i=i0;i<=i1;i++)for(int j=j0;j<=j1;j++){ if(owner[i*W+j]!=r) continue; long di=i-ci,dj=j-cj,d2=di*di+dj*dj,dot=di*p0+dj*p1;
            if(!found||d2<bd2||(d2==bd2&&dot>bdot)){found=1;bdi=di;bdj=dj;bd2=d2;bdot=dot;alt=0;} else if(d2==bd2&&dot==bdot){alt=1;adi=di;adj=dj;} }
        }
        rad>>=1;
        steps++; long c=labs(bdi)>labs(bdj)?labs(bdi):labs(bdj); dist[c==1?0:c==2?1:c<=4?2:c<=8?3:c<=16?4:5]++;
        int pin5=labs(p0)<=2&&labs(p1)<=2; prefin5+=pin5; if(c<=2){found5++; found5_pref5+=pin5;}
        if(alt){ tie++; int ex=exact_pref(p0,p1);
          double nrm=sqrt((double)bd2); double sb=fma((double)bdj/nrm,(double)p1,((double)bdi/nrm)*(double)p0), sa=fma((double)adj/nrm,(double)p1,((double)adi/nrm)*(double)p0);
          long cb=bdi*p1-bdj*p0, ca=adi*p1-adj*p0; int ab = sa!=sb? sa>sb : ca>cb;
          if(!ex){ tie_ne++; long m=labs(p0); if(labs(p1)>m)m=labs(p1); if(labs(bdi)>m)m=labs(bdi); if(labs(bdj)>m)m=labs(bdj); if(labs(adi)>m)m=labs(adi); if(labs(adj)>m)m=labs(adj);
            tie_ne_r[m<=3?0:m<=7?1:m<=15?2:3]++; if(adi==-bdi&&adj==-bdj) tie_opp++;
            if((labs(adi)==labs(bdi)&&labs(adj)==labs(bdj))||(labs(adi)==labs(bdj)&&labs(adj)==labs(bdi))) tie_sameset++;
            tie_res[sa==sb?0:(sa>sb?1:2)]++; }
          if(ab){bdi=adi;bdj=adj;} }
        int bi=ci+bdi,bj=cj+bdj; owner[bi*W+bj]=-1; path[a+t]=bi*W+bj; p0=bdi;p1=bdj;ci=bi;cj=bj;
      }
    }
    // reduce
    int k=0; for(int r=0;r<R;r++){ noff[r]=k; for(int g=off[r];g<off[r+1];g++) if((g&1)==0) next[k++]=path[g]; } noff[R]=k;
    printf("lev %2d steps %7ld cheb1 %5.1f%% 2 %5.1f%% 3-4 %5.1f%% 5-8 %5.1f%% 9-16 %4.1f%% >16 %4.1f%% | pref in5x5 %5.1f%% found5 %5.1f%% both %5.1f%% | tie %5.1f%% nonexact %5.1f%% (opp %ld sameset %ld; m<=3 %ld <=7 %ld <=15 %ld >15 %ld; eq %ld a>b %ld a<b %ld)\n",lev,steps,
      100.*dist[0]/steps,100.*dist[1]/steps,100.*dist[2]/steps,100.*dist[3]/steps,100.*dist[4]/steps,100.*dist[5]/steps,100.*prefin5/steps,100.*found5/steps,100.*found5_pref5/steps,100.*tie/steps,100.*tie_ne/steps,tie_opp,tie_sameset,tie_ne_r[0],tie_ne_r[1],tie_ne_r[2],tie_ne_r[3],tie_res[0],tie_res[1],tie_res[2]);
    tot_steps+=steps; tot_tie+=tie; tot_tie_nonexact+=tie_ne; for(int q=0;q<4;q++) tot_inrange[q]+=tie_ne_r[q];
    if(lev>1) free(cur); cur=next; free(off); off=noff; n_l=k; free(path);
    if(n_l<2) break;
  }
  printf("total steps %ld ties %ld (%.1f%%) non-exact %ld (%.1f%%) ranges %ld %ld %ld %ld\n",tot_steps,tot_tie,100.*tot_tie/tot_steps,tot_tie_nonexact,100.*tot_tie_nonexact/tot_steps,tot_inrange[0],tot_inrange[1],tot_inrange[2],tot_inrange[3]);
  return 0; }
