for v in auto 2,2 3,2; do echo "=== SFV_EPI_SLOTS=$v"; SFV_EPI_SLOTS=$v python tools/layer_profile.py --one | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('JSON::'):
        d=json.loads(l[6:]); print('total', round(d['total_ms'],2)); 
        for i,r in enumerate(d['recs'][:50]):
            if r['cat']==0 and 'res=1' in r['tag']: print(i, round(r['ms']*1e3,1), r['tag'])
"; done
