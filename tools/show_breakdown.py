import json,sys
for f in sys.argv[1:]:
    d=json.load(open(f))
    print(f, 'warps',d.get('warps'),'rpb',d.get('rpb'),'cap',d.get('table_slots'), d["amp_bead_mean"], {k:round(v,2) for k,v in d["all"].items()})
    print('   ', ' '.join(f'{k[:6]}={d[k]["us_per_attempt"]:.2f}us/{d[k]["share"]:.2f}' for k in ["crank_shaft","end_pivot","slide","tangent_rotation","change_binding_state"]), 'sum_sweep_us', round(d['sum_sweep_us'],1))
