"""Write the CSV logs of three small fixtures with the UNMODIFIED reference's own CSVLogger
(free_range_zoo/utils/logging_handlers.py:36-111) into tests/golden/logs/<fixture>/<env>.csv.

Run in the build container only (needs /root/reference):  python tests/golden/gen_logs.py

The rollouts are the seeded ones of gen_golden.py (same actions, same injected uniforms as the .npz fixtures of the
same name), only with ``log_directory`` set.  tests/test_logging_gpu.py replays those fixtures on the engine with its
asynchronous logging tap and compares the files byte for byte.
"""
import os
import shutil

import gen_golden as G

from free_range_zoo_b200 import presets

G.save = lambda *args, **kwargs: None  # the .npz fixtures are not rewritten
ROOT = os.path.join(G.HERE, 'logs')
KEEP = 4  # environments whose files are committed (1 for the large grid)


def run(name, generate, *args, keep=KEEP, **kwargs):
    scratch = os.path.join('/tmp', f'frz_logs_{name}')
    shutil.rmtree(scratch, ignore_errors=True)
    generate(name, *args, log_directory=scratch, **kwargs)
    target = os.path.join(ROOT, name)
    shutil.rmtree(target, ignore_errors=True)
    os.makedirs(target)
    for env in range(keep):
        shutil.copy(os.path.join(scratch, f'{env}.csv'), os.path.join(target, f'{env}.csv'))
    print(name, sum(os.path.getsize(os.path.join(target, f)) for f in os.listdir(target)), 'bytes')


if __name__ == '__main__':
    run('wildfire_profile', G.gen_wildfire, presets.wildfire_profile, B=16, steps=15, seed=11)
    run('wildfire_c4', G.gen_wildfire, presets.wildfire_large, B=8, steps=30, seed=13, keep=1)
    run('rideshare_profile', G.gen_rideshare, presets.rideshare_profile, B=8, steps=20, seed=21)
    run('cyber_profile', G.gen_cyber, presets.cyber_profile, B=8, steps=20, seed=31)
    run('cyber_c3', G.gen_cyber, presets.cyber_c3, B=32, steps=60, seed=32, show_bad_actions=False, partially_observable=True)
