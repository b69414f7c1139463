import sys; sys.path.insert(0,'/root/repo')
from mri_image_generation_b200 import _lib
l=_lib.load()
for bn,st in [(128,3),(128,4),(128,6),(64,4),(16,5),(256,4),(256,2)]:
    print('block_n',bn,'stages',st,'smem',l.mri_gemm_smem_bytes(bn,st),'occupancy',l.mri_gemm_occupancy(bn,st))
