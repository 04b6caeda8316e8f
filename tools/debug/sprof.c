#define _GNU_SOURCE
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <ucontext.h>
#include <unistd.h>
#define NB (1<<20)
static unsigned long buf[NB]; static volatile unsigned long n;
static void handler(int s, siginfo_t *si, void *uc){ unsigned long i=__sync_fetch_and_add(&n,1); if(i<NB) buf[i]=((ucontext_t*)uc)->uc_mcontext.gregs[REG_RIP]; }
__attribute__((constructor)) static void init(void){ struct sigaction sa; memset(&sa,0,sizeof sa); sa.sa_sigaction=handler; sa.sa_flags=SA_SIGINFO|SA_RESTART; sigaction(SIGPROF,&sa,0);
  struct itimerval it={{0,500},{0,500}}; setitimer(ITIMER_PROF,&it,0); }
__attribute__((destructor)) static void fini(void){ struct itimerval it={{0,0},{0,0}}; setitimer(ITIMER_PROF,&it,0);
  const char *out=getenv("SPROF_OUT"); if(!out) out="/tmp/sprof.out"; FILE*f=fopen(out,"w"); FILE*m=fopen("/proc/self/maps","r"); char line[512];
  while(fgets(line,sizeof line,m)) if(strstr(line,"libk")||strstr(line,"libm")||strstr(line,"libc")) fprintf(f,"M %s",line); fclose(m);
  unsigned long k=n<NB?n:NB; for(unsigned long i=0;i<k;i++) fprintf(f,"%lx\n",buf[i]); fclose(f); }
