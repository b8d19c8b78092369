// Chat client for the OpenAI-compatible endpoint of this server: streaming fetch + SSE parsing,
// settings and named sessions in localStorage, regenerate / edit-and-resend, tokens/s read-out.
(function () {
  "use strict";
  var $ = function (id) { return document.getElementById(id); };
  var KEY_CFG = "b200chat.settings", KEY_SESS = "b200chat.sessions";
  var cfgFields = { endpoint: "cfg-endpoint", apikey: "cfg-apikey", model: "cfg-model", stop: "cfg-stop",
                    maxtokens: "cfg-maxtokens", temperature: "cfg-temperature", topp: "cfg-topp" };
  var sessions = JSON.parse(localStorage.getItem(KEY_SESS) || "{}");
  var current = null, controller = null;

  function loadCfg() {
    var c = JSON.parse(localStorage.getItem(KEY_CFG) || "{}");
    Object.keys(cfgFields).forEach(function (k) { if (c[k] !== undefined) $(cfgFields[k]).value = c[k]; });
  }
  function readCfg() {
    var c = {};
    Object.keys(cfgFields).forEach(function (k) { c[k] = $(cfgFields[k]).value; });
    return c;
  }
  function saveSessions() { localStorage.setItem(KEY_SESS, JSON.stringify(sessions)); }
  function newSession() {
    var id = "s" + Date.now();
    sessions[id] = { title: "New chat", messages: [] };
    current = id; saveSessions(); renderSessions(); renderMessages();
  }
  function renderSessions() {
    var ul = $("session-list"); ul.innerHTML = "";
    Object.keys(sessions).sort().reverse().forEach(function (id) {
      var li = document.createElement("li");
      li.textContent = sessions[id].title; li.className = id === current ? "active" : "";
      li.onclick = function () { current = id; renderSessions(); renderMessages(); };
      var del = document.createElement("button"); del.textContent = "×"; del.title = "Delete";
      del.onclick = function (e) { e.stopPropagation(); delete sessions[id]; if (current === id) current = null; saveSessions(); if (!Object.keys(sessions).length) newSession(); else { current = current || Object.keys(sessions)[0]; renderSessions(); renderMessages(); } };
      li.appendChild(del); ul.appendChild(li);
    });
  }
  function bubble(msg, idx) {
    var div = document.createElement("div"); div.className = "msg " + msg.role;
    var body = document.createElement("div"); body.className = "body";
    body.innerHTML = msg.role === "assistant" ? window.renderMarkdown(msg.content) : "";
    if (msg.role !== "assistant") body.textContent = msg.content;
    div.appendChild(body);
    var tools = document.createElement("div"); tools.className = "tools";
    if (msg.role === "assistant") {
      var re = document.createElement("button"); re.textContent = "Regenerate";
      re.onclick = function () { sessions[current].messages.splice(idx); saveSessions(); renderMessages(); generate(); };
      tools.appendChild(re);
    } else {
      var ed = document.createElement("button"); ed.textContent = "Edit";
      ed.onclick = function () { $("prompt").value = msg.content; sessions[current].messages.splice(idx); saveSessions(); renderMessages(); $("prompt").focus(); };
      tools.appendChild(ed);
    }
    div.appendChild(tools);
    return div;
  }
  function renderMessages() {
    var box = $("messages"); box.innerHTML = "";
    (sessions[current] ? sessions[current].messages : []).forEach(function (m, i) { box.appendChild(bubble(m, i)); });
    box.scrollTop = box.scrollHeight;
  }
  function setBusy(b) { $("send").hidden = b; $("stop").hidden = !b; }

  async function generate() {
    var c = readCfg(), sess = sessions[current];
    var payload = { model: c.model || "default_model", messages: sess.messages.map(function (m) { return { role: m.role, content: m.content }; }),
                    stream: true, max_tokens: parseInt(c.maxtokens || "512", 10), temperature: parseFloat(c.temperature || "0.7"),
                    top_p: parseFloat(c.topp || "0.95") };
    if (c.stop) payload.stop = [c.stop];
    var headers = { "Content-Type": "application/json" };
    if (c.apikey) headers.Authorization = "Bearer " + c.apikey;
    var reply = { role: "assistant", content: "" };
    sess.messages.push(reply); renderMessages();
    var bodyEl = $("messages").lastChild.querySelector(".body");
    controller = new AbortController(); setBusy(true);
    var t0 = performance.now(), tFirst = null, chunks = 0;
    try {
      var resp = await fetch(c.endpoint || "/v1/chat/completions", { method: "POST", headers: headers, body: JSON.stringify(payload), signal: controller.signal });
      if (!resp.ok) throw new Error("HTTP " + resp.status + ": " + (await resp.text()));
      var reader = resp.body.getReader(), dec = new TextDecoder(), buf = "";
      for (;;) {
        var r = await reader.read(); if (r.done) break;
        buf += dec.decode(r.value, { stream: true });
        var parts = buf.split("\n\n"); buf = parts.pop();
        parts.forEach(function (evt) {
          evt.split("\n").forEach(function (line) {
            if (line.indexOf("data:") !== 0) return;
            var data = line.slice(5).trim();
            if (!data || data === "[DONE]") return;
            var obj = JSON.parse(data);
            if (obj.error) throw new Error(obj.error.message);
            var ch = obj.choices && obj.choices[0];
            var piece = ch && ((ch.delta && ch.delta.content) || ch.text || "");
            if (piece) { if (tFirst === null) tFirst = performance.now(); chunks++; reply.content += piece; bodyEl.innerHTML = window.renderMarkdown(reply.content); $("messages").scrollTop = $("messages").scrollHeight; }
          });
        });
      }
    } catch (e) {
      if (e.name !== "AbortError") { reply.content += "\n\n*[error: " + e.message + "]*"; bodyEl.innerHTML = window.renderMarkdown(reply.content); }
    } finally {
      controller = null; setBusy(false);
      if (sess.title === "New chat" && sess.messages.length) sess.title = sess.messages[0].content.slice(0, 32);
      saveSessions(); renderSessions();
      if (tFirst !== null) $("stats").textContent = "TTFT " + (tFirst - t0).toFixed(0) + " ms · " + (chunks / Math.max((performance.now() - tFirst) / 1000, 1e-3)).toFixed(1) + " chunks/s";
    }
  }

  $("composer").addEventListener("submit", function (e) {
    e.preventDefault();
    var text = $("prompt").value.trim(); if (!text || controller) return;
    sessions[current].messages.push({ role: "user", content: text }); $("prompt").value = "";
    saveSessions(); renderMessages(); generate();
  });
  $("prompt").addEventListener("keydown", function (e) { if (e.key === "Enter" && !e.shiftKey) { e.preventDefault(); $("composer").requestSubmit(); } });
  $("stop").onclick = function () { if (controller) controller.abort(); };
  $("new-session").onclick = newSession;
  $("cfg-save").onclick = function (e) { e.preventDefault(); localStorage.setItem(KEY_CFG, JSON.stringify(readCfg())); $("settings").open = false; };
  loadCfg();
  if (!Object.keys(sessions).length) newSession(); else { current = Object.keys(sessions).sort().reverse()[0]; renderSessions(); renderMessages(); }
})();
