// reader_bench.cpp -- throughput of the host ingest path (mfkc_reader_*) on FASTA/FASTQ(.gz) files, best of 5 passes.
//   g++ -O2 -o /tmp/reader_bench tools/reader_bench.cpp -Lmetafast_b200/lib -lmfkc -Wl,-rpath,$PWD/metafast_b200/lib
//   MFKC_INFLATE=zlib|serial, MFKC_INFLATE_THREADS=N, MFKC_READER_THREADS=N select the older paths / thread counts
#include "../include/mfkc.h"
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <vector>
static double now(){struct timespec t;clock_gettime(CLOCK_MONOTONIC,&t);return t.tv_sec+t.tv_nsec*1e-9;}
int main(int argc,char**argv){
  const size_t cap_bases=256u<<20; const uint32_t cap_reads=1u<<21;
  std::vector<uint8_t> bases(cap_bases); std::vector<uint64_t> offs(cap_reads+1);
  for(size_t i=0;i<cap_bases;i+=4096)bases[i]=1;
  for(int a=1;a<argc;a++){ double best=1e9; unsigned long long tot=0;
    for(int rep=0;rep<5;rep++){ mfkc_reader*r=nullptr; char err[256]; if(mfkc_reader_open(argv[a],&r,err,sizeof err)){printf("%s\n",err);return 1;}
      double t=now(); tot=0; for(;;){uint32_t n=0; if(mfkc_reader_next(r,bases.data(),cap_bases,offs.data(),cap_reads,&n)){printf("err %s\n",mfkc_reader_error(r));break;} if(!n)break; tot+=n;}
      double dt=now()-t; if(dt<best)best=dt; mfkc_reader_close(r);}
    printf("%s: %llu reads best %.3f s  %.2f M reads/s\n",argv[a],tot,best,tot/best/1e6);}
}
