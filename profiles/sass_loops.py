"""Static SASS loop statistics of the fp32 K=3 sweep kernel family in a built library:
python profiles/sass_loops.py hmc.jl_b200/lib/libhmcgpu.so   (instruction counts per loop body, spills, memory ops)"""
import re, subprocess, sys
out = subprocess.run(['cuobjdump', '-sass', sys.argv[1]], capture_output=True, text=True).stdout
funcs, cur = {}, None
for l in out.splitlines():
    m = re.search(r'Function : (\S+)', l)
    if m:
        cur = m.group(1); funcs[cur] = []
    elif cur and re.search(r'/\*[0-9a-f]{4,}\*/\s+\S', l):
        funcs[cur].append(l)
pat = sys.argv[2] if len(sys.argv) > 2 else 'IfLi3ELb0ELb0ELb0'
addr = lambda l: int(re.search(r'/\*([0-9a-f]+)\*/', l).group(1), 16)
for name, L in funcs.items():
    if pat not in name:
        continue
    print(name[:110], len(L), 'instr')
    for l in L:
        m = re.search(r'BRA(?:\.U)? .*?(0x[0-9a-f]+)', l)
        if m and int(m.group(1), 16) < addr(l):
            body = [x for x in L if int(m.group(1), 16) <= addr(x) <= addr(l)]
            c = lambda k: sum(k in x for x in body)
            if 40 < len(body) < 1200:
                print(f"   loop {m.group(1)}..{addr(l):#x}: {len(body)} instr, spill {c('LDL') + c('STL')}, LDG {c('LDG')}, LDS {c('LDS')}, "
                      f"LDGSTS {c('LDGSTS')}, EX2 {c('MUFU.EX2')}, STG {c('STG')}, philox-mul {c('-0x2daee0ad')}")
