'''
`yanocomp` command group (reference yanocomp/cli.py): `prep` and `gmmtest`.
click_log is not a dependency here, so the `-v/--verbosity` option it
provides is re-created by hand.
'''
import logging

import click

from . import dist as ycdist
from .gmmtest import gmm_test
from .prep import nanopolish_collapse


def get_logger(name):
    logger = logging.getLogger(name)
    if not logger.handlers:
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter('{asctime} {levelname:8s} {message}', style='{'))
        logger.addHandler(handler)
        logger.propagate = False
    return logger


logger = get_logger('yanocomp')


def _set_verbosity(ctx, param, value):
    level = getattr(logging, str(value).upper(), None)
    if not isinstance(level, int):
        raise click.BadParameter('Must be CRITICAL, ERROR, WARNING, INFO or DEBUG, not {}'.format(value))
    logger.setLevel(level)
    return value


def _log_profile(prof, top=50):
    import io
    import pstats
    prof.disable()
    buf = io.StringIO()
    pstats.Stats(prof, stream=buf).sort_stats('cumulative').print_stats(top)
    logger.debug('Profiling complete\n%s', buf.getvalue())


def _profile_until_exit():
    '''hidden --profile flag (reference cli.py:36-50): host-side cProfile over the whole process, the 50 entries
    with the largest cumulative time logged at DEBUG level on exit; the kernels are profiled with ncu instead'''
    import atexit
    import cProfile
    prof = cProfile.Profile()
    atexit.register(_log_profile, prof)
    logger.debug('Running with profiler...')
    prof.enable()


@click.command(short_help='Convert a prep output (HDF5 / packed) to the dense sample container gmmtest ingests fastest')
@click.argument('in_fn', type=click.Path(exists=True, dir_okay=False))
@click.argument('out_fn', type=click.Path(dir_okay=False))
@click.option('-p', '--processes', default=1, show_default=True, type=click.IntRange(1, None),
              help='Host threads decoding genes side by side')
def pack(in_fn, out_fn, processes):
    '''
    IN_FN: one sample's collapsed events (the reference's HDF5 or this package's packed prep output).
    OUT_FN: the dense container (yanocomp_b200/dense.py): the per-file half of gmmtest's pivot done once.
    Not in the reference; `gmmtest -c / -t` take either kind of file.
    '''
    from . import dense
    n = [0]

    def tick(_):
        n[0] += 1
    dense.convert(in_fn, out_fn, progress=tick, threads=processes)
    logger.info(f'{n[0]:,} genes written to {out_fn}')


@click.group(commands={'prep': nanopolish_collapse, 'gmmtest': gmm_test, 'pack': pack})
@click.option('-v', '--verbosity', default='INFO', metavar='LVL', expose_value=False,
              callback=_set_verbosity, is_eager=True,
              help='Either CRITICAL, ERROR, WARNING, INFO or DEBUG')
@click.option('--profile', is_flag=True, hidden=True)
def cli(profile):
    if profile:
        _profile_until_exit()
    world, _ = ycdist.init_from_env()
    if world > 1:
        import atexit
        import torch.distributed as dist
        atexit.register(lambda: dist.is_initialized() and dist.destroy_process_group())


if __name__ == '__main__':
    cli()
