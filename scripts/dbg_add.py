import sys, faulthandler, numpy as np
faulthandler.enable()
sys.path.insert(0, '.')
from tests import harness as H
from colli2de_b200 import scenes
sc = scenes.mixed_scene(30000, 3000, 900.0)
for nb in (0, 1):
    if nb: sc["ent_body"][::40] = 2
    for chunk in (5000, 20000, 33000):
        b = H.Backend("b200")
        b.create_entities(sc["ent_id"][:chunk], sc["ent_body"][:chunk], sc["ent_xf3"][:chunk])
        print("created", nb, chunk, flush=True)
        ids = b.add_shapes(sc["shp_entity"][:chunk], sc["shp_type"][:chunk], sc["shp_geom"][:chunk], sc["shp_count"][:chunk], sc["shp_cat"][:chunk], sc["shp_mask"][:chunk])
        print("added", nb, chunk, flush=True)
        p = b.get_colliding_pairs()
        print("pairs", len(p), flush=True)
