"""Device time of the one-pass fractional-yaw kernel (p2p_project_views_table) on a resident 8192x4096 panorama: one yaw and
four yaws (one coordinate evaluation shared by the four) x three pitches of the README example.
    python tools/time_fractional_yaw.py
"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.build(); pkg = g.load_package()
from tools import synth_inputs as synth
import torch
proj = pkg.Projector(0, n_slots=2)
WP, HP, W, H = 8192, 4096, 1920, 1080
consts = [pkg.pitch_constants(W, 100, p) for p in (30, 60, 90)]
tabs = [pkg.yaw_table(WP, y)[:2] for y in (33.3, 123.3, 213.3, 303.3)]
d = torch.empty((4, 3, H, W, 3), dtype=torch.uint8, device='cuda:0')
with proj.slots(1) as (s,):
    proj.upload(s, synth.noise(WP, HP, 0)); proj.sync(s)
    ev0, ev1 = proj.event(), proj.event()
    for n in (1, 4):
        for _ in range(3):
            proj.record(ev0, s); proj.project_tables(s, tabs[:n], consts, W, H, out_device_ptr=d.data_ptr()); proj.record(ev1, s); proj.sync(s)
        print(n, 'yaws x 3 pitches:', round(proj.elapsed_ms(ev0, ev1) * 1e3, 1), 'us')
