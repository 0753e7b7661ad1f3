"""Synthetic syslog-style corpus and the flume-style rule sets used by the
tests and by bench.py (BASELINE.json configs[0..2]).

The rules are modelled on the reference's own flume configuration
(quol/flume/cfg.tsv: `sshd.*Failed password for`, `smtpd.*lost connection`,
...).  A union of 64 two-part `prog.*phrase` rules is not compilable by RED
itself (the subset construction explodes: 16 of them already give 6904 states
and take 37 s in the reference), so the 64-rule set is what a RED user would
actually load: the 64 distinguishing phrases, each added with fLooseStart and
its own result id, matched with styInstant ("contained": the line contains the
phrase; end = offset just past it).

Everything is deterministic in the seed.
"""

import numpy as np

F_IGNORE_CASE, F_LOOSE_START, F_LOOSE_END = 1, 2, 4
LANG_RAW, LANG_AUTO, LANG_GLOB, LANG_EXACT = 1, 2, 3, 4

PHRASES = [
    "Accepted password for", "Accepted publickey for", "Failed password for root",
    "Failed password for invalid user", "Failed password for", "Connection closed by",
    "Did not receive identification string", "Received disconnect from", "Bad protocol",
    "Unable to negotiate", ": login:", "action=greylist", "action=pass", "Passed CLEAN",
    "Quarantined", "TLS negotiation failed", "Connection reset by peer", "Connection timed out",
    ": Delivered:", "exception", "status=sent", ": client=", "lost connection",
    "does not resolve to address", "Host not found", "cannot find your hostname",
    "Domain not found", "blocked using zen.spamhaus.org", "blocked using bl.spamcop.net",
    "blocked using b.barracudacentral.org", "too many errors", "Relay access denied",
    "User unknown in local recipient table", "status=deferred", "status=bounced",
    "SSL_accept error", "SASL LOGIN authentication", "SASL PLAIN authentication",
    "SASL authentication failure", "non-SMTP command from", "Sensor alarm", "DHCPDISCOVER",
    "DHCPREQUEST", "DHCPACK", "/etc/passwd", "segfault at", "Out of memory", "oom-killer",
    "I/O error", "link is down", "link is up", "authentication failure",
    "session opened for user", "session closed for user", "POSSIBLE BREAK-IN ATTEMPT",
    "Invalid user", "error: maximum authentication attempts exceeded",
    "kernel: nf_conntrack: table full", "temperature above threshold", "EXT4-fs error",
    "UFW BLOCK", "Started Session", "Stopping", "denied",
]
assert len(PHRASES) == 64


def rules_single():
    """configs[0]: one syslog regex, grep-style (addAuto: loose at both ends)."""
    return [("sshd.*Failed password", 1, 0, LANG_AUTO)]


def rules_union64():
    """configs[1]: 64 rules, distinct result ids, loose start."""
    return [(p, i + 1, F_LOOSE_START, LANG_RAW) for i, p in enumerate(PHRASES)]


def rules_keywords(n, seed=7, min_len=5, max_len=12):
    """configs[2]: n exact keywords (addExact, loose start) -> a large
    Aho-Corasick-like DFA that does not fit shared memory."""
    rng = np.random.RandomState(seed)
    seen = set()
    out = []
    alphabet = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz0123456789_-", dtype=np.uint8)
    while len(out) < n:
        ln = rng.randint(min_len, max_len + 1)
        w = bytes(alphabet[rng.randint(0, len(alphabet), ln)])
        if w in seen:
            continue
        seen.add(w)
        out.append((w, len(out) + 1, F_LOOSE_START, LANG_EXACT))
    return out


_HOSTS = ["web%02d" % i for i in range(24)] + ["mx%d" % i for i in range(4)] + ["db%d" % i for i in range(6)]
_USERS = ["root", "admin", "deploy", "alice", "bob", "carol", "oracle", "postgres", "git", "www-data"]
_MONTHS = ["Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"]


def _ip(r):
    return "%d.%d.%d.%d" % (r.randint(1, 224), r.randint(0, 256), r.randint(0, 256), r.randint(1, 255))


def _qid(r):
    return "".join("0123456789ABCDEF"[x] for x in r.randint(0, 16, 10))


def _templates():
    # (weight, function(rng) -> message after the "host prog[pid]: " prefix)
    T = []
    A = T.append
    # noise: matches none of the 64 phrases
    A((14, "CRON", lambda r: "(%s) CMD (/usr/local/bin/rotate --quiet --keep %d)" % (_USERS[r.randint(10)], r.randint(1, 30))))
    A((12, "systemd", lambda r: "Starting Daily apt download activities for unit %d..." % r.randint(1000)))
    A((10, "kernel", lambda r: "[%d.%06d] audit: type=1400 apparmor=\"ALLOWED\" operation=\"open\" pid=%d" % (r.randint(1, 999999), r.randint(0, 999999), r.randint(100, 65000))))
    A((10, "nginx", lambda r: "%s - - \"GET /api/v%d/items/%d HTTP/1.1\" 200 %d \"-\" \"curl/8.%d\"" % (_ip(r), r.randint(1, 4), r.randint(1, 100000), r.randint(200, 90000), r.randint(0, 9))))
    A((8, "postfix/qmgr", lambda r: "%s: from=<%s@example.org>, size=%d, nrcpt=1 (queue active)" % (_qid(r), _USERS[r.randint(10)], r.randint(400, 90000))))
    A((8, "dbus-daemon", lambda r: "[system] Activating via systemd: service name='org.freedesktop.hostname%d'" % r.randint(9)))
    A((7, "ntpd", lambda r: "adjusting local clock by %d.%06ds" % (r.randint(0, 3), r.randint(0, 999999))))
    A((6, "postfix/cleanup", lambda r: "%s: message-id=<%s.%d@mail.example.org>" % (_qid(r), _qid(r), r.randint(1, 99999))))
    # lines that match (phrase id in comment is 1-based result)
    A((3, "sshd", lambda r: "Failed password for invalid user %s from %s port %d ssh2" % (_USERS[r.randint(10)], _ip(r), r.randint(1024, 65535))))
    A((2, "sshd", lambda r: "Failed password for root from %s port %d ssh2" % (_ip(r), r.randint(1024, 65535))))
    A((2, "sshd", lambda r: "Accepted publickey for %s from %s port %d ssh2: RSA SHA256:%s" % (_USERS[r.randint(10)], _ip(r), r.randint(1024, 65535), _qid(r))))
    A((2, "sshd", lambda r: "Connection closed by %s port %d [preauth]" % (_ip(r), r.randint(1024, 65535))))
    A((1, "sshd", lambda r: "Received disconnect from %s port %d:11: Bye Bye [preauth]" % (_ip(r), r.randint(1024, 65535))))
    A((2, "postfix/smtp", lambda r: "%s: to=<%s@example.com>, relay=mx.example.com[%s]:25, delay=%d.%d, status=sent (250 ok)" % (_qid(r), _USERS[r.randint(10)], _ip(r), r.randint(0, 30), r.randint(0, 99))))
    A((2, "postfix/smtpd", lambda r: "connect from unknown[%s]: client=%s lost connection after AUTH" % (_ip(r), _ip(r))))
    A((1, "postfix/smtpd", lambda r: "NOQUEUE: reject: RCPT from unknown[%s]: 554 5.7.1 Service unavailable; blocked using zen.spamhaus.org" % _ip(r)))
    A((1, "postgrey", lambda r: "action=greylist, reason=new, client_name=unknown, client_address=%s" % _ip(r)))
    A((1, "amavis", lambda r: "(%05d-%02d) Passed CLEAN {RelayedInbound}, [%s] <%s@example.net>" % (r.randint(0, 99999), r.randint(0, 99), _ip(r), _USERS[r.randint(10)])))
    A((1, "kernel", lambda r: "[%d.%06d] [UFW BLOCK] IN=eth0 OUT= SRC=%s DST=%s PROTO=TCP DPT=%d" % (r.randint(1, 999999), r.randint(0, 999999), _ip(r), _ip(r), r.randint(1, 65535))))
    A((1, "dhcpd", lambda r: "DHCPACK on %s to %02x:%02x:%02x:%02x:%02x:%02x via eth%d" % ((_ip(r),) + tuple(r.randint(0, 256, 6)) + (r.randint(0, 4),))))
    A((1, "systemd-logind", lambda r: "pam_unix(sshd:session): session opened for user %s by (uid=%d)" % (_USERS[r.randint(10)], r.randint(0, 2000))))
    A((1, "kernel", lambda r: "[%d.%06d] EXT4-fs error (device sd%s%d): ext4_find_entry:%d: inode #%d" % (r.randint(1, 999999), r.randint(0, 999999), "abcd"[r.randint(4)], r.randint(1, 9), r.randint(1000, 2000), r.randint(2, 1 << 20))))
    return T


def make_lines(n, seed=1234):
    """n synthetic syslog lines (bytes, no newline)."""
    r = np.random.RandomState(seed)
    T = _templates()
    w = np.array([t[0] for t in T], dtype=np.float64)
    w /= w.sum()
    picks = r.choice(len(T), size=n, p=w)
    out = []
    for k in picks:
        _, prog, fn = T[k]
        ts = "%s %2d %02d:%02d:%02d" % (_MONTHS[r.randint(12)], r.randint(1, 29), r.randint(24), r.randint(60), r.randint(60))
        out.append(("%s %s %s[%d]: %s" % (ts, _HOSTS[r.randint(len(_HOSTS))], prog, r.randint(100, 99999), fn(r))).encode("ascii"))
    return out


def pack_lines(lines, newline=True):
    """-> (text uint8, starts uint64, lens uint32); lines are laid end to end,
    separated by '\\n' when newline is set (the separator is not in lens)."""
    lens = np.fromiter((len(x) for x in lines), dtype=np.uint32, count=len(lines))
    step = lens.astype(np.uint64) + (1 if newline else 0)
    starts = np.zeros(len(lines), dtype=np.uint64)
    if len(lines) > 1:
        np.cumsum(step[:-1], out=starts[1:])
    sep = b"\n" if newline else b""
    blob = sep.join(lines) + sep
    text = np.frombuffer(blob, dtype=np.uint8)
    return text, starts, lens


def make_corpus(n, seed=1234, newline=True):
    return pack_lines(make_lines(n, seed), newline)
