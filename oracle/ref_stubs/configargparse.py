import argparse
class ArgumentParser(argparse.ArgumentParser):
    def add_argument(self, *a, is_config_file=False, **k):
        return super().add_argument(*a, **k)
    def _cfg_to_argv(self, path):
        out = []
        for line in open(path):
            line = line.split('#')[0].strip()
            if not line or '=' not in line: continue
            k, v = [s.strip() for s in line.split('=', 1)]
            if v == 'True': out.append('--' + k)
            elif v == 'False': continue
            elif v.startswith('['): out += ['--' + k] + [s.strip() for s in v.strip('[]').split(',')]
            else: out += ['--' + k, v]
        return out
    def parse_known_args(self, args=None, namespace=None):
        import sys
        args = list(sys.argv[1:] if args is None else args)
        if '--config' in args:
            i = args.index('--config'); args = self._cfg_to_argv(args[i+1]) + args[:i] + args[i+2:]
        return super().parse_known_args(args, namespace)
