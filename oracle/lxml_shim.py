"""A stand-in for the part of ``lxml.etree`` that ``aku2elan.py`` uses (aku2elan.py:5, 45-99).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

lxml is not installed in this image, so ``oracle/ref_exec.py`` registers this module as
``lxml.etree`` while it executes the reference's ``aku2elan.py``: the reference's own code then
builds the element tree (every tag, attribute and value is the reference's), and only the
SERIALISATION below is ours.  It follows what lxml / libxml2 write for
``ElementTree.write(file, pretty_print=True)`` under Python 2:

* no XML declaration (the default encoding is ASCII; non-ASCII characters become ``&#N;``);
* attributes given as a ``dict`` are stored sorted by name (lxml's ``_iter_attrib`` on
  interpreters without ordered dicts) - a Clark-notation name ``{ns}local`` sorts after the
  upper-case names; namespace declarations come first in the start tag, the XML-Schema-instance
  namespace gets lxml's default prefix ``xsi``;
* two-space indentation, ``<tag/>`` for empty elements, text-only elements on one line,
  one trailing newline.

PARITY NOTE: this serialisation is restated from lxml's documented behaviour, not pinned by a
run of lxml (absent here, version un-pinned in the reference's Dockerfile)."""

_DEFAULT_PREFIX = {'http://www.w3.org/2001/XMLSchema-instance': 'xsi',
                   'http://www.w3.org/XML/1998/namespace': 'xml'}


class _Element(object):
    def __init__(self, tag, attrib=None):
        self.tag = tag
        self.attrib = sorted((attrib or {}).items())
        for k, v in self.attrib:
            if not isinstance(v, str):
                raise TypeError('Argument must be bytes or unicode, got %r' % type(v).__name__)
        self.text = None
        self.children = []


def Element(tag, attrib=None):
    return _Element(tag, attrib)


def SubElement(parent, tag, attrib=None):
    e = _Element(tag, attrib)
    parent.children.append(e)
    return e


def _esc(s, attr):
    out = []
    for ch in s:
        if ch == '&':
            out.append('&amp;')
        elif ch == '<':
            out.append('&lt;')
        elif ch == '>':
            out.append('&gt;')
        elif attr and ch == '"':
            out.append('&quot;')
        elif attr and ch in '\n\r\t':
            out.append('&#%d;' % ord(ch))
        elif ord(ch) > 126:
            out.append('&#%d;' % ord(ch))
        else:
            out.append(ch)
    return ''.join(out)


def _start(e, root):
    ns = []
    attrs = []
    for k, v in e.attrib:
        if k.startswith('{'):
            uri, local = k[1:].split('}')
            prefix = _DEFAULT_PREFIX.get(uri, 'ns%d' % len(ns))
            if (prefix, uri) not in ns:
                ns.append((prefix, uri))
            k = prefix + ':' + local
        attrs.append(' %s="%s"' % (k, _esc(v, True)))
    decl = ''.join(' xmlns:%s="%s"' % (p, _esc(u, True)) for p, u in ns)
    return '<' + e.tag + decl + ''.join(attrs)


def _write(e, depth, out):
    pad = '  ' * depth
    head = _start(e, depth == 0)
    if not e.children and e.text is None:
        out.append(pad + head + '/>\n')
    elif not e.children:
        out.append(pad + head + '>' + _esc(e.text, False) + '</' + e.tag + '>\n')
    else:
        # libxml2 indents children only when the element has no text of its own
        out.append(pad + head + '>\n')
        for c in e.children:
            _write(c, depth + 1, out)
        out.append(pad + '</' + e.tag + '>\n')


class ElementTree(object):
    def __init__(self, root):
        self.root = root

    def write(self, outf, pretty_print=False):
        assert pretty_print, 'only the pretty-printed form is restated'
        out = []
        _write(self.root, 0, out)
        text = ''.join(out)
        if hasattr(outf, 'write'):
            outf.write(text)
        else:
            with open(outf, 'w') as f:
                f.write(text)
