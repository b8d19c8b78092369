// Tiny dependency-free markdown renderer (the GPU box is offline, so nothing is pulled from a CDN):
// fenced code, inline code, headings, bold/italic, links, lists, paragraphs.  Output is escaped first.
(function (global) {
  function esc(s) { return s.replace(/&/g, "&amp;").replace(/</g, "&lt;").replace(/>/g, "&gt;"); }
  function inline(s) {
    return s
      .replace(/`([^`]+)`/g, function (_, c) { return "<code>" + c + "</code>"; })
      .replace(/\*\*([^*]+)\*\*/g, "<strong>$1</strong>")
      .replace(/(^|[^*])\*([^*]+)\*/g, "$1<em>$2</em>")
      .replace(/\[([^\]]+)\]\((https?:[^)\s]+)\)/g, '<a href="$2" target="_blank" rel="noopener">$1</a>');
  }
  function render(src) {
    var out = [], lines = esc(src).split("\n"), i = 0, para = [], list = null;
    function flushPara() { if (para.length) { out.push("<p>" + inline(para.join("<br>")) + "</p>"); para = []; } }
    function flushList() { if (list) { out.push("<" + list.tag + ">" + list.items.map(function (t) { return "<li>" + inline(t) + "</li>"; }).join("") + "</" + list.tag + ">"); list = null; } }
    while (i < lines.length) {
      var ln = lines[i];
      var fence = ln.match(/^```(\w*)\s*$/);
      if (fence) {
        flushPara(); flushList();
        var code = []; i++;
        while (i < lines.length && !/^```\s*$/.test(lines[i])) { code.push(lines[i]); i++; }
        out.push('<pre><code class="lang-' + fence[1] + '">' + code.join("\n") + "</code></pre>");
        i++; continue;
      }
      var h = ln.match(/^(#{1,6})\s+(.*)$/);
      if (h) { flushPara(); flushList(); out.push("<h" + h[1].length + ">" + inline(h[2]) + "</h" + h[1].length + ">"); i++; continue; }
      var ul = ln.match(/^\s*[-*]\s+(.*)$/), ol = ln.match(/^\s*\d+[.)]\s+(.*)$/);
      if (ul || ol) {
        flushPara();
        var tag = ul ? "ul" : "ol";
        if (!list || list.tag !== tag) { flushList(); list = { tag: tag, items: [] }; }
        list.items.push((ul || ol)[1]); i++; continue;
      }
      if (/^\s*$/.test(ln)) { flushPara(); flushList(); i++; continue; }
      flushList(); para.push(ln); i++;
    }
    flushPara(); flushList();
    return out.join("\n");
  }
  global.renderMarkdown = render;
})(window);
