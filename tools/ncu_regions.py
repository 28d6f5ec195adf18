import subprocess,sys,re,collections
out=subprocess.run(['python','tools/ncu_lines.py',sys.argv[1],sys.argv[2],'400'],capture_output=True,text=True).stdout.splitlines()
src=open('rasterizer_b200/csrc/orz_kernels.cu').read().splitlines()
def find(pat,start=0):
    for i,l in enumerate(src[start:],start):
        if pat in l: return i+1
    return None
marks=[('raster_preamble',find('__device__ __forceinline__ void raster_prim')),('row_loop',find('for (uint32_t by = 0; by < rangeY')),('segment',find('for (uint32_t s0 = a; s0 < b')),('cand_loop',find('while (cand) {')),('chain_steps',find('for (uint32_t i = 0; i < steps')),('post_steps',find('owed = 0; pos = j;')),('convex_mask',find('if (convex) {  // Rasterizer')),('nonconvex_mask',find('} else {  // Rasterizer.cpp:1188')),('update',find('uint32_t* dptr = depthWords')),('hiz',find('uint32_t mn = min(val')),('row_end',find('owed += m - pos;')),('after_raster',find('// query2D, Rasterizer.cpp:283-349')),('query',find('block_fine_test')),('setup_chunk',find('void setup_chunk')),('frame',find('k_render_views(const FrameParams p)')),('end',find('k_query_views(const FrameParams'))]
agg=collections.Counter(); smp=collections.Counter()
for l in out[3:]:
    m=re.match(r'(\S+):(\d+)\s+([\d.]+)\s+([\d.]+)',l)
    if not m: continue
    f,n,i,s=m.group(1),int(m.group(2)),float(m.group(3)),float(m.group(4))
    if f!='orz_kernels.cu': reg='other:'+f
    else:
        reg='pre'
        for name,ln in marks:
            if ln and n>=ln: reg=name
    agg[reg]+=i; smp[reg]+=s
for k,v in sorted(agg.items(),key=lambda kv:-kv[1]): print(f'{k:40s} inst {v:6.2f}%  samples {smp[k]:6.2f}%')
