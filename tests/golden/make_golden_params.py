"""Seeds / shapes shared by make_golden.py and the tests that re-create the
fixture inputs (inputs are regenerated from seeds, never stored)."""
from riser_b200 import synth

POLYA_SEED, POLYA_READS = 31, 40
POLYA_PREFIXES = [400, 500, 999, 1500, 2500, 4000, 5500, 6000, 7500, 9000, 12000, 15000,
                  18528, 18529, 20000]
NORM_SEED, NORM_READS = 2024, 24
NORM_EXTRA_LENGTHS = [4096, 4097, 7108, 8615, 10120, 12048, 12047, 5000]
SCEN = dict(seed=99, n_reads=40, kit="RNA002", chunk=3012, first_len=6024, n_polls=6,
            targets=["mRNA", "mtRNA"], threshold=0.9)


def polya_reads():
    return synth.raw_reads(POLYA_SEED, POLYA_READS, min_body=3000, max_body=16000,
                           frac_no_polya=0.15, frac_const=0.05)


def norm_inputs():
    """Ragged body set shared by the normalise and ConvNet goldens."""
    bodies = synth.ragged_bodies(NORM_SEED, NORM_READS, 4096, 12048)
    extra = synth.body_batch(NORM_SEED + 1, len(NORM_EXTRA_LENGTHS), 12048)
    bodies += [extra[i, :n].copy() for i, n in enumerate(NORM_EXTRA_LENGTHS)]
    return bodies


def scenario_reads():
    return synth.raw_reads(SCEN["seed"], SCEN["n_reads"], min_body=3000, max_body=15000,
                           frac_no_polya=0.15, frac_const=0.05)
