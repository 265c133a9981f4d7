#!/usr/bin/env python
"""Per CUDA line of one source file: warp instructions, stall samples, threads per instruction, with inlined helpers folded
into the nearest preceding line of that file (address order).  Usage: ncu_bylines.py dump.csv file.cuh path/to/file.cuh [min_pct]"""
import csv, sys
dump, fname, path = sys.argv[1:4]
minp = float(sys.argv[4]) if len(sys.argv) > 4 else 0.6
rows=[];cur=None;ln=None;ix=None
for row in csv.reader(open(dump,newline='')):
    if not row: continue
    if row[0]=="File Path": cur=row[1].split('/')[-1]; continue
    if row[0]=="Line No": ix={n:i for i,n in enumerate(row)}; continue
    if row[0]=="Function Name" or ix is None: continue
    if row[0]!="": ln=int(row[0]); continue
    if row[2]=="...": continue
    try: inst=int(float(row[ix["Instructions Executed"]])); smp=int(float(row[ix["# Samples"]])); thr=int(float(row[ix["Thread Instructions Executed"]]))
    except ValueError: continue
    rows.append((int(row[2],16),cur,ln,inst,smp,thr))
by={}
for a,f,l,i,s,t in rows:
    if a not in by or (f==fname and by[a][0]!=fname): by[a]=(f,l,i,s,t)
tot=sum(v[2] for v in by.values()); tots=sum(v[3] for v in by.values()) or 1
agg={};last=0
for a in sorted(by):
    f,l,i,s,t=by[a]
    if f==fname: last=l
    x=agg.setdefault(last,[0,0,0]); x[0]+=i; x[1]+=s; x[2]+=t
src=open(path).read().split('\n')
print("warp instructions", tot, "samples", tots)
for l in sorted(agg):
    if 100*agg[l][0]/tot>=minp or 100*agg[l][1]/tots>=2*minp:
        print(f"{l:4d} {100*agg[l][0]/tot:5.2f}% inst {100*agg[l][1]/tots:5.1f}% smp thr/inst {agg[l][2]/max(agg[l][0],1):4.1f} {src[l-1].strip()[:95] if l else ''}")
