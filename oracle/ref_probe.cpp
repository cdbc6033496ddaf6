#include "NucleotideMatrix.h"
#include "EvalueComputation.h"
#include "LocalParameters.h"
#include <cstdio>
// stand-alone Sls dump
int main(){ LocalParameters::getLocalInstance();
  std::string nm="nucleotide.out"; std::string md; { FILE*f=fopen("/root/reference/lib/mmseqs/data/nucleotide.out","rb"); char b[4096]; size_t n=fread(b,1,4096,f); md.assign(b,n); fclose(f);}
  char* ser=BaseMatrix::serialize(nm,md);
  NucleotideMatrix subMat(ser, 1.0, 0.0);
  printf("alphabetSize %d name %s\n", subMat.alphabetSize, subMat.getMatrixName().c_str());
  for(int i=0;i<subMat.alphabetSize;i++) printf("pBack[%d]=%.20g\n",i,subMat.pBack[i]);
  Sls::AlignmentEvaluer ev;
  long *tmp[5]; long d[25];
  for(int i=0;i<5;i++){tmp[i]=&d[i*5]; for(int j=0;j<5;j++) d[i*5+j]=subMat.subMatrix[i][j];}
  ev.initGapless(4,(const long*const*)tmp,subMat.pBack,subMat.pBack,60.0);
  const Sls::ALP_set_of_parameters &p=ev.parameters();
  printf("lambda %.20g\nK %.20g\nC %.20g\na_I %.20g\nb_I %.20g\nalpha_I %.20g\nbeta_I %.20g\na_J %.20g\nb_J %.20g\nalpha_J %.20g\nbeta_J %.20g\nsigma %.20g\ntau %.20g\nvi_y_thr %.20g\nvj_y_thr %.20g\nc_y_thr %.20g\n",
   p.lambda,p.K,p.C,p.a_I,p.b_I,p.alpha_I,p.beta_I,p.a_J,p.b_J,p.alpha_J,p.beta_J,p.sigma,p.tau,p.vi_y_thr,p.vj_y_thr,p.c_y_thr);
  EvalueComputation e(766712, &subMat);
  int scores[]={0,1,10,34,35,38,40,60,96,140,240,2000};
  int qls[]={20,35,48,70,150,1167,200000};
  size_t dbs[]={766712, 10500000, 525000000ull, 45000000000ull};
  for(size_t dbr: dbs){ EvalueComputation e2(dbr,&subMat);
   for(int s:scores) for(int q:qls){
    double ev_=e2.computeEvalue(s,q); double bs=e2.computeBitScore(s);
    printf("KAT %zu %d %d %.17g %.17g\n",dbr,s,q,ev_,bs);}}
  return 0;
}
