"""
HOCON-subset reader for the reference's ``conf/*.conf`` model/renderer schema.

The reference reads its configs with pyhocon (``src/util/args.py:6,90-101``) and then
queries the tree with ``get_int / get_float / get_bool / get_string / get_list``,
``conf["model"]`` and dotted keys (``resnetfc.py:238-250``, ``encoder.py:235-252``,
``code.py:49-56``, ``nerf.py:340-352``).  pyhocon is not a dependency of this package;
this module implements exactly the grammar the shipped confs use:

* ``key = value`` / ``key : value`` / ``key { ... }`` (nested objects merge recursively,
  later scalars override earlier ones),
* ``include required("relative/path.conf")`` and ``include "path"``,
* ``#`` and ``//`` comments, optional commas,
* scalars: ``true/false/True/False``, ints, floats (incl. ``5e-4``), quoted or bare strings,
  ``null``; lists ``[a, b, [c, d]]``.

Unknown keys are kept and ignored by the consumers (``conf/exp/sn64_multiscale.conf``
carries ``fusion_*``, ``hard_alpha_cap`` ... that only the broken live model reads).
"""
import os
import re

__all__ = ["ConfigTree", "ConfigFactory", "ConfigMissing"]

_MISSING = object()


class ConfigMissing(KeyError):
    pass


class ConfigTree(dict):
    """Nested dict with the pyhocon accessor surface the reference uses."""

    # ---- lookup ----------------------------------------------------------
    def _find(self, key):
        node = self
        for part in str(key).split("."):
            if not isinstance(node, dict) or not dict.__contains__(node, part):
                return _MISSING
            node = dict.__getitem__(node, part)
        return node

    def get(self, key, default=_MISSING):
        val = self._find(key)
        if val is _MISSING:
            if default is _MISSING:
                raise ConfigMissing("No configuration setting found for key %s" % key)
            return default
        return val

    def __getitem__(self, key):
        val = self._find(key)
        if val is _MISSING:
            raise ConfigMissing("No configuration setting found for key %s" % key)
        return val

    def __contains__(self, key):
        return self._find(key) is not _MISSING

    def _typed(self, key, default, conv):
        val = self.get(key, default)
        if val is None or val is default:
            return val
        return conv(val)

    def get_int(self, key, default=_MISSING):
        return self._typed(key, default, lambda v: int(v))

    def get_float(self, key, default=_MISSING):
        return self._typed(key, default, lambda v: float(v))

    def get_string(self, key, default=_MISSING):
        def conv(v):
            if isinstance(v, bool):
                return "true" if v else "false"
            return str(v)

        return self._typed(key, default, conv)

    def get_bool(self, key, default=_MISSING):
        def conv(v):
            if isinstance(v, str):
                low = v.strip().lower()
                if low in ("true", "yes", "on"):
                    return True
                if low in ("false", "no", "off"):
                    return False
                raise ValueError("%s is not a boolean: %r" % (key, v))
            return bool(v)

        return self._typed(key, default, conv)

    def get_list(self, key, default=_MISSING):
        def conv(v):
            if not isinstance(v, list):
                raise ValueError("%s is not a list: %r" % (key, v))
            return v

        return self._typed(key, default, conv)

    def get_config(self, key, default=_MISSING):
        return self.get(key, default)

    # ---- construction ----------------------------------------------------
    def put(self, key, value):
        parts = str(key).split(".")
        node = self
        for part in parts[:-1]:
            nxt = dict.get(node, part)
            if not isinstance(nxt, ConfigTree):
                nxt = ConfigTree()
                dict.__setitem__(node, part, nxt)
            node = nxt
        last = parts[-1]
        old = dict.get(node, last)
        if isinstance(old, ConfigTree) and isinstance(value, ConfigTree):
            old.merge(value)
        else:
            dict.__setitem__(node, last, value)

    def merge(self, other):
        for k, v in other.items():
            self.put_local(k, v)
        return self

    def put_local(self, key, value):
        old = dict.get(self, key)
        if isinstance(old, ConfigTree) and isinstance(value, ConfigTree):
            old.merge(value)
        elif isinstance(value, ConfigTree):
            dict.__setitem__(self, key, ConfigTree().merge(value))
        else:
            dict.__setitem__(self, key, value)

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, ConfigTree) else v) for k, v in self.items()}


_TOKEN_RE = re.compile(
    r"""
    (?P<ws>[ \t\r]+)
  | (?P<nl>\n)
  | (?P<comment>(\#|//)[^\n]*)
  | (?P<string>"(?:\\.|[^"\\])*")
  | (?P<punct>[{}\[\],=:])
  | (?P<bare>[^\s{}\[\],=:"\#]+)
    """,
    re.VERBOSE,
)

_INT_RE = re.compile(r"^[+-]?\d+$")
_FLOAT_RE = re.compile(r"^[+-]?(\d+\.\d*|\.\d+|\d+)([eE][+-]?\d+)?$")


def _scalar(text):
    low = text.lower()
    if low == "true":
        return True
    if low == "false":
        return False
    if low == "null":
        return None
    if _INT_RE.match(text):
        return int(text)
    if _FLOAT_RE.match(text):
        return float(text)
    return text


class _Parser:
    def __init__(self, text, basedir):
        self.toks = []
        pos = 0
        while pos < len(text):
            m = _TOKEN_RE.match(text, pos)
            if m is None:
                raise ValueError("conf: cannot tokenize at %r" % text[pos : pos + 20])
            pos = m.end()
            kind = m.lastgroup
            if kind in ("ws", "comment"):
                continue
            self.toks.append((kind, m.group(kind)))
        self.i = 0
        self.basedir = basedir

    def peek(self):
        return self.toks[self.i] if self.i < len(self.toks) else ("eof", "")

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def skip_sep(self):
        while self.peek()[0] == "nl" or self.peek() == ("punct", ","):
            self.i += 1

    def parse_object(self, closing):
        tree = ConfigTree()
        while True:
            self.skip_sep()
            kind, val = self.peek()
            if kind == "eof":
                if closing:
                    raise ValueError("conf: missing '}'")
                return tree
            if (kind, val) == ("punct", "}"):
                if not closing:
                    raise ValueError("conf: stray '}'")
                self.next()
                return tree
            if kind == "bare" and val == "include":
                self.next()
                tree.merge(self.parse_include())
                continue
            if kind == "bare":
                key = val
            elif kind == "string":
                key = val[1:-1]
            else:
                raise ValueError("conf: expected key, got %r" % (val,))
            self.next()
            kind, val = self.peek()
            if (kind, val) == ("punct", "{"):
                self.next()
                tree.put(key, self.parse_object(True))
                continue
            if kind == "punct" and val in "=:":
                self.next()
                tree.put(key, self.parse_value())
                continue
            raise ValueError("conf: expected '=', ':' or '{' after key %r" % key)

    def parse_include(self):
        kind, val = self.next()
        required = False
        if kind == "bare" and val.startswith("required"):
            # tokenised as  required("path")  ->  bare 'required(' ... handle both forms
            rest = val[len("required") :]
            required = True
            if rest.startswith("("):
                rest = rest[1:]
            if rest:
                path = rest
            else:
                kind, val = self.next()
                path = val
            path = path.strip('()"')
            # swallow a trailing ')' token if the string was separate
            if self.peek()[0] == "bare" and self.peek()[1] == ")":
                self.next()
        elif kind == "string":
            path = val[1:-1]
        else:
            raise ValueError("conf: bad include %r" % (val,))
        full = path if os.path.isabs(path) else os.path.join(self.basedir, path)
        if not os.path.exists(full):
            if required:
                raise FileNotFoundError("conf: required include not found: %s" % full)
            return ConfigTree()
        return ConfigFactory.parse_file(full)

    def parse_value(self):
        kind, val = self.next()
        if (kind, val) == ("punct", "{"):
            return self.parse_object(True)
        if (kind, val) == ("punct", "["):
            return self.parse_list()
        if kind == "string":
            return bytes(val[1:-1], "utf-8").decode("unicode_escape")
        if kind == "bare":
            # bare values may contain spaces up to end of line: join consecutive bare tokens
            parts = [val]
            while self.peek()[0] == "bare":
                parts.append(self.next()[1])
            if len(parts) == 1:
                return _scalar(val)
            return " ".join(parts)
        raise ValueError("conf: bad value %r" % (val,))

    def parse_list(self):
        out = []
        while True:
            self.skip_sep()
            kind, val = self.peek()
            if (kind, val) == ("punct", "]"):
                self.next()
                return out
            if kind == "eof":
                raise ValueError("conf: missing ']'")
            out.append(self.parse_value())


class ConfigFactory:
    """``ConfigFactory.parse_file`` / ``parse_string`` as used at ``src/util/args.py:90-101``."""

    @staticmethod
    def parse_string(text, basedir="."):
        # 'include required("x")' tokenises awkwardly because of the parentheses; normalise first.
        text = re.sub(r'include\s+required\(\s*"([^"]*)"\s*\)', r'include required("\1")', text)
        text = re.sub(r'required\("([^"]*)"\)', r'required "\1"', text)
        return _Parser(text, basedir).parse_object(False)

    @staticmethod
    def parse_file(path):
        with open(path, "r", encoding="utf-8") as fh:
            text = fh.read()
        return ConfigFactory.parse_string(text, os.path.dirname(os.path.abspath(path)))

    @staticmethod
    def from_dict(d):
        tree = ConfigTree()
        for k, v in d.items():
            tree.put_local(k, ConfigFactory.from_dict(v) if isinstance(v, dict) else v)
        return tree
