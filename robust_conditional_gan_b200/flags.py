"""The reference's flag surface without TensorFlow: a minimal stand-in for `tf.app.flags` (mnist/main.py:12-67,
cifar10/gan_resnet.py:38-76) with the same command-line syntax -- `--name value`, `--name=value`, and for booleans
`--name` / `--noname` / `--name=false` -- so the run scripts (mnist/run_*.sh, cifar10/run_*.sh) drive this package unchanged.
Unknown flags raise, as absl does."""
import sys


class FlagValues(object):
    def __init__(self):
        self.__dict__['_defs'] = {}
        self.__dict__['_vals'] = {}

    def _define(self, kind, name, default, doc):
        self._defs[name] = (kind, doc)
        self._vals[name] = default

    def __getattr__(self, name):
        if name == '__flags':            # the reference prints `flags.FLAGS.__flags` (mnist/main.py:83)
            return dict(self.__dict__['_vals'])
        try:
            return self.__dict__['_vals'][name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self._vals[name] = value

    def __contains__(self, name):
        return name in self._vals

    def flag_values_dict(self):
        return dict(self._vals)

    def _convert(self, name, text):
        kind = self._defs[name][0]
        if kind == 'integer':
            return int(text)
        if kind == 'float':
            return float(text)
        if kind == 'boolean':
            if text.lower() in ('true', 't', '1', 'yes'):
                return True
            if text.lower() in ('false', 'f', '0', 'no'):
                return False
            raise ValueError('flag --%s: %r is not a boolean' % (name, text))
        if kind == 'list':
            return [t for t in text.split(',') if t != '']
        return text

    def parse(self, argv):
        """Parses argv (without the program name) in place of absl; returns the positional leftovers."""
        rest, i = [], 0
        while i < len(argv):
            a = argv[i]
            i += 1
            if a == '--':
                rest += argv[i:]
                break
            if not a.startswith('-') or a == '-':
                rest.append(a)
                continue
            body = a.lstrip('-')
            name, eq, val = body.partition('=')
            if name in self._defs:
                if self._defs[name][0] == 'boolean' and not eq:
                    self._vals[name] = True
                    continue
                if not eq:
                    if i >= len(argv):
                        raise ValueError('flag --%s needs a value' % name)
                    val = argv[i]
                    i += 1
                self._vals[name] = self._convert(name, val)
            elif name.startswith('no') and name[2:] in self._defs and self._defs[name[2:]][0] == 'boolean' and not eq:
                self._vals[name[2:]] = False
            else:
                raise ValueError('Unknown command line flag %r' % name)
        return rest


class Flags(object):
    """`flags = Flags()` plays the role of `tf.app.flags`."""

    def __init__(self):
        self.FLAGS = FlagValues()

    def DEFINE_integer(self, name, default, doc=''):
        self.FLAGS._define('integer', name, default, doc)

    def DEFINE_float(self, name, default, doc=''):
        self.FLAGS._define('float', name, default, doc)

    def DEFINE_string(self, name, default, doc=''):
        self.FLAGS._define('string', name, default, doc)

    def DEFINE_boolean(self, name, default, doc=''):
        self.FLAGS._define('boolean', name, default, doc)

    DEFINE_bool = DEFINE_boolean

    def DEFINE_list(self, name, default, doc=''):
        self.FLAGS._define('list', name, default, doc)


def run(main, flags, argv=None):
    """tf.app.run(): parse the command line, then main(argv)."""
    argv = sys.argv if argv is None else argv
    rest = flags.FLAGS.parse(list(argv[1:]))
    return main([argv[0]] + rest)
