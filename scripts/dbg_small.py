import sys; sys.path.insert(0,'scripts')
import gpu_check as g
for M,P in [(8,8),(16,8),(64,64),(128,128)]:
    g.check(M,P,backend='direct')
