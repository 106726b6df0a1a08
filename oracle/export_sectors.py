#!/usr/bin/env python
"""Export the reference's static sector DATA (scenarios.py:14-32 SimpleScenario, :35-207 LOWW) to the
sector file format of this repo (JSON, see atc_reinforcement_learning_b200/sector.py).  Data only — polygon
vertices in the reference's order, MVA heights, runway, entry points.  Build container only.

    python oracle/export_sectors.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'standins'))
sys.path.insert(1, os.environ.get('ATC_REFERENCE_ROOT', '/root/reference'))
from envs.atc import scenarios  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'atc_reinforcement_learning_b200', 'sectors')


def eps(scn):
    return [{'x': float(e.x), 'y': float(e.y), 'phi': float(e.phi), 'levels': [int(l) for l in e.levels]}
            for e in scn.entrypoints]


def export(name):
    cls = getattr(scenarios, name)
    scn, scn_r = cls(random_entrypoints=False), cls(random_entrypoints=True)
    doc = {
        'format': 'atc-b200-sector/1',
        'name': name,
        'source': 'fvalka/atc-reinforcement-learning envs/atc/scenarios.py',
        'units': {'xy': 'nm', 'height': 'ft', 'phi': 'deg compass', 'levels': 'flight level (x100 ft)'},
        # vertex order and MVA order are significant: first match wins (model.py:282-289)
        'mvas': [{'height': int(m.height), 'ring': [[float(x), float(y)] for x, y in m.area.exterior.coords]}
                 for m in scn.mvas],
        'runway': {'x': float(scn.runway.x), 'y': float(scn.runway.y), 'h': float(scn.runway.h),
                   'phi_from_runway': float(scn.runway.phi_from_runway)},
        'entrypoints': eps(scn),
        'entrypoints_random': eps(scn_r),
    }
    with open(os.path.join(OUT, name + '.json'), 'w') as f:
        json.dump(doc, f, indent=1)
    print(name, len(doc['mvas']), 'mvas', sum(len(m['ring']) for m in doc['mvas']), 'ring vertices')


if __name__ == '__main__':
    export('LOWW')
    export('SimpleScenario')
