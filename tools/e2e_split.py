import time, numpy as np, sys
sys.path.insert(0,'/root/repo')
from go_gsgp_b200 import Engine, capi
from go_gsgp_b200.host import HostDriver
from bench import synth
P,V,N=1000,20,100000
X,y=synth(2,N,V); nrow=80000
host=HostDriver(population_size=P,nvar=V,seed=1); sym=host.create_T_F(); trees=host.initialize_population(0)
e=Engine(population_size=P,nvar=V,nrow=nrow,nrow_test=N-nrow,error_measure=capi.RMSE)
e.set_dataset(X,y); e.set_constants(sym); e.eval_initial(0,trees); fit,ft,best=e.fitness_all()
tb=tg=0
for i in range(60):
    t0=time.perf_counter(); plan,pool=host.build_plan(fit,best); t1=time.perf_counter()
    fit,ft,best=e.generation(plan,pool); t2=time.perf_counter()
    if i>=10: tb+=t1-t0; tg+=t2-t1
print('build_plan %.3f ms  generation() %.3f ms'%(tb/50*1e3,tg/50*1e3))
e.profile_enable(True)
for i in range(20):
    plan,pool=host.build_plan(fit,best); fit,ft,best=e.generation(plan,pool)
n,ms,_=e.profile_read(capi.K_GSOP); print('gen kernel ms', ms/n)
