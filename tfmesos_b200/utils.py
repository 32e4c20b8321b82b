"""Control-plane wire format and small helpers.

Wire format is the reference's (tfmesos/utils.py:6-15): a 4-byte big-endian
length followed by a pickle.  Unlike the reference, ``recv`` loops until the
whole frame has arrived (the reference does a single ``fd.recv(size)``, which
short-reads once a config dict outgrows one TCP segment -- SURVEY.md 4.3), and
``send`` uses ``sendall``.  Frames are byte-compatible in both directions.
"""
import logging
import pickle
import socket
import struct

_LEN = struct.Struct('>I')


def send(fd, o):
    payload = pickle.dumps(o)
    fd.sendall(_LEN.pack(len(payload)) + payload)


def _read_exact(fd, n):
    chunks = []
    while n > 0:
        chunk = fd.recv(n)
        if not chunk:
            break
        chunks.append(chunk)
        n -= len(chunk)
    return b''.join(chunks)


def recv(fd):
    head = _read_exact(fd, _LEN.size)
    assert len(head) == _LEN.size, repr(head)
    size, = _LEN.unpack(head)
    body = _read_exact(fd, size)
    assert len(body) == size, 'short frame: %d of %d bytes' % (len(body), size)
    return pickle.loads(body)


def setup_logger(logger):
    """Same record format as tfmesos/utils.py:18-27."""
    fmt = logging.Formatter('%(asctime)-11s [%(levelname)s] [%(name)-9s] %(message)s')
    logger.setLevel(logging.INFO)
    handler = logging.StreamHandler()
    handler.setLevel(logging.DEBUG)
    handler.setFormatter(fmt)
    logger.addHandler(handler)
    return logger


class AttrDict(dict):
    """Attribute-style nested dict used for offers, TaskInfos and status
    updates (the reference gets this from the third-party ``addict``)."""

    def __init__(self, *args, **kw):
        dict.__init__(self)
        for a in args:
            for k, v in dict(a).items():
                self[k] = v
        for k, v in kw.items():
            self[k] = v

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        try:
            return self[name]
        except KeyError:
            node = self[name] = AttrDict()
            return node

    def __setattr__(self, name, value):
        self[name] = value


def local_hostname():
    """Host part of every advertised address.  The reference uses
    socket.gethostname() (scheduler.py:328, server.py:21); on a single box that
    name may not resolve, in which case the loopback address is used."""
    name = socket.gethostname()
    try:
        socket.getaddrinfo(name, None)
        return name
    except socket.error:
        return '127.0.0.1'


def bind_advertised(sock, port=0, host=None):
    """Bind ``sock`` to the address this box ADVERTISES (``local_hostname()``, or the
    host part of a cluster-spec address) instead of every interface: the control
    sockets unpickle what they receive and the task endpoint runs what its client
    asks for, so nothing off the box has any business connecting.  The
    reference binds ``''`` (server.py:18-21, scheduler.py:328-333); set
    ``TFMESOS_BIND=all`` for that.  Falls back to ``''`` if the advertised name
    cannot be bound here (a hostname that resolves to a foreign address)."""
    import os
    host = local_hostname() if host is None else host
    if os.environ.get('TFMESOS_BIND', 'advertised') != 'all':
        try:
            sock.bind((host, port))
            return
        except OSError:
            pass
    sock.bind(('', port))
