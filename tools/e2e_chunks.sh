for c in 8 6 5 4; do
  METRPO_E2E_CHUNKS=$c timeout 200 python bench.py --steps 6 2>/dev/null > /tmp/b_$c.json
  python - <<P
import json
d=json.loads(open('/tmp/b_$c.json').read().strip().splitlines()[-1])
print("chunks", $c, round(d["value"]/1e6,1), round(d["e2e"]["value"]/1e6,1), d["clocks"]["sm_mhz"])
P
done

