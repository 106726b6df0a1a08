"""Sector definitions (mirror of /root/reference/envs/atc/scenarios.py:7-207).  The polygon / runway / entry-point
DATA lives in sectors/*.json (format 'atc-b200-sector/1', exported from the reference by oracle/export_sectors.py);
user sectors can be loaded from the same format with `load_scenario(path)`."""
import json
import os
from typing import List

from . import model

SECTOR_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'sectors')


class Scenario(object):
    """scenarios.py:7-11"""
    name: str
    runway: model.Runway
    mvas: List[model.MinimumVectoringAltitude]
    entrypoints: List[model.EntryPoint]

    def __init__(self, doc=None, random_entrypoints=False):
        if doc is not None:
            self._from_doc(doc, random_entrypoints)

    def _from_doc(self, doc, random_entrypoints):
        if doc.get('format') != 'atc-b200-sector/1':
            raise ValueError("unknown sector file format %r" % doc.get('format'))
        self.name = doc['name']
        self.mvas = [model.MinimumVectoringAltitude(m['ring'], m['height']) for m in doc['mvas']]
        if not self.mvas:
            raise ValueError("sector has no MVA polygons")
        r = doc['runway']
        self.runway = model.Runway(r['x'], r['y'], r['h'], r['phi_from_runway'])
        eps = doc.get('entrypoints_random') if random_entrypoints else doc.get('entrypoints')
        if not eps:
            raise ValueError("sector has no entry points")
        self.entrypoints = [model.EntryPoint(e['x'], e['y'], e['phi'], e['levels']) for e in eps]
        self.random_entrypoints = bool(random_entrypoints)


def load_scenario(path, random_entrypoints=False):
    with open(path) as f:
        return Scenario(json.load(f), random_entrypoints)


class SimpleScenario(Scenario):
    """scenarios.py:14-32"""

    def __init__(self, random_entrypoints=False):
        with open(os.path.join(SECTOR_DIR, 'SimpleScenario.json')) as f:
            super().__init__(json.load(f), random_entrypoints)


class LOWW(Scenario):
    """scenarios.py:35-207 — Vienna approach: 12 MVA polygons, runway (45.16, 43.26, 586 ft, 160), 1 or 9 entry points"""

    def __init__(self, random_entrypoints=False):
        with open(os.path.join(SECTOR_DIR, 'LOWW.json')) as f:
            super().__init__(json.load(f), random_entrypoints)
