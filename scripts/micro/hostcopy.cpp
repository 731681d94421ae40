#include <thread>
#include <vector>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include <cstdint>
#include <algorithm>
#include <cstdlib>
static double now(){return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();}
int main(int argc,char**argv){
  const uint64_t N=512ull<<20; const char*path="/dev/shm/rr.bin";
  { std::vector<uint8_t> d(N, 7); int fd=open(path,O_CREAT|O_WRONLY|O_TRUNC,0644); if(write(fd,d.data(),N)<0) return 1; close(fd);}
  int fd=open(path,O_RDONLY);
  uint8_t*buf=(uint8_t*)aligned_alloc(4096, N); memset(buf,1,N);
  // (a) T threads, each preads a contiguous 1/T of the whole file in 8 MiB calls
  for (unsigned T : {1u,2u,4u,8u}) {
    double t0=now();
    std::vector<std::thread> th;
    for(unsigned t=0;t<T;t++) th.emplace_back([&,t]{ uint64_t lo=N/T*t, hi=N/T*(t+1); for(uint64_t a=lo;a<hi;a+=8<<20){ uint64_t n=std::min<uint64_t>(8<<20,hi-a); uint64_t d=0; while(d<n){ssize_t g=pread(fd,buf+a+d,n-d,a+d); if(g<=0)break; d+=g;} } });
    for(auto&x:th)x.join();
    double dt=now()-t0; printf("pread whole-file, %u threads: %.1f ms %.2f GB/s\n", T, dt*1e3, N/dt/1e9);
  }
  // (b) mmap + memcpy
  for (int populate : {0,1}) for (unsigned T : {1u,4u,8u}) {
    double t0=now();
    uint8_t*m=(uint8_t*)mmap(nullptr,N,PROT_READ,MAP_SHARED|(populate?MAP_POPULATE:0),fd,0);
    double t1=now();
    std::vector<std::thread> th;
    for(unsigned t=0;t<T;t++) th.emplace_back([&,t]{ uint64_t lo=N/T*t, hi=N/T*(t+1); memcpy(buf+lo,m+lo,hi-lo); });
    for(auto&x:th)x.join();
    double dt=now()-t0; printf("mmap(populate=%d) %.1f ms + memcpy %u threads: total %.1f ms %.2f GB/s\n", populate,(t1-t0)*1e3, T, dt*1e3, N/dt/1e9);
    munmap(m,N);
  }
  // memcpy baseline
  { uint8_t*src=(uint8_t*)aligned_alloc(4096,N); memset(src,3,N); for (unsigned T : {1u,4u,8u}) { double t0=now(); std::vector<std::thread> th;
    for(unsigned t=0;t<T;t++) th.emplace_back([&,t]{ uint64_t lo=N/T*t, hi=N/T*(t+1); memcpy(buf+lo,src+lo,hi-lo); }); for(auto&x:th)x.join(); double dt=now()-t0; printf("plain memcpy %u threads: %.1f ms %.2f GB/s\n",T,dt*1e3,N/dt/1e9);} }
  unlink(path);
}
