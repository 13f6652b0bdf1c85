"""Print a C++ source file without comments, keeping original line numbers (study aid)."""
import sys
s=open(sys.argv[1],encoding='utf-8').read()
i=0;n=len(s);line=1;res=[];cur='';state=0
while i<n:
    c=s[i]
    if state==0:
        if s.startswith('//',i): state=1;i+=2;continue
        if s.startswith('/*',i): state=2;i+=2;continue
        if c=='"':
            j=i+1
            while s[j]!='"' or s[j-1]=='\\': j+=1
            cur+=s[i:j+1]; i=j+1; continue
        if c=='\n': res.append((line,cur));cur='';line+=1
        else: cur+=c
    elif state==1:
        if c=='\n': state=0;res.append((line,cur));cur='';line+=1
    else:
        if s.startswith('*/',i): state=0;i+=2;continue
        if c=='\n': res.append((line,cur));cur='';line+=1
    i+=1
for l,t in res:
    if t.strip(): print(f"{l}: {t.rstrip()}")
